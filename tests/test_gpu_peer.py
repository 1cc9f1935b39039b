"""The peer-memory all-reduce of the per-step exchange (csrc/peer.cu).

World size 1 runs on the single test GPU (the kernel's push / flag / sum protocol against its own area, many epochs);
world size 2 needs two GPUs and is skipped otherwise (scripts/peer_check.py under torchrun: bit-identity with the
rank-ordered sum and agreement with NCCL, 200 epochs)."""
import ctypes
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_world_of_one_is_the_identity_over_many_epochs():
    from pde_read_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    cap = 4096
    area, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
    assert lib.pderead_peer_area_bytes(1, cap) == 512 + 2 * cap * 4
    _lib.check(lib.pderead_peer_alloc(1, cap, ctypes.byref(area), handle), "pderead_peer_alloc")
    try:
        table = torch.tensor([area.value], dtype=torch.int64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        g = torch.Generator(device=dev).manual_seed(5)
        for n in (4, 1000, 4096, 8, 4092) * 5:                      # odd and even epochs, sizes up to the capacity
            x = torch.randn(n, device=dev, generator=g)
            want = x.clone()
            _lib.check(lib.pderead_peer_allreduce(x.data_ptr(), n, table.data_ptr(), area.value, 0, 1, cap, stream),
                       "pderead_peer_allreduce")
            assert torch.equal(x, want)
        # argument checks: length not a multiple of 4, above the capacity, bad rank
        x = torch.zeros(8192, device=dev)
        assert lib.pderead_peer_allreduce(x.data_ptr(), 6, table.data_ptr(), area.value, 0, 1, cap, stream) != 0
        assert lib.pderead_peer_allreduce(x.data_ptr(), 8192, table.data_ptr(), area.value, 0, 1, cap, stream) != 0
        assert lib.pderead_peer_allreduce(x.data_ptr(), 8, table.data_ptr(), area.value, 1, 1, cap, stream) != 0
        torch.cuda.synchronize()
    finally:
        lib.pderead_peer_free(area)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_match_nccl_and_the_rank_ordered_sum():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29577", os.path.join(ROOT, "scripts", "peer_check.py")]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "peer ok: True" in p.stdout and "mismatches over 200 epochs (all ranks): 0" in p.stdout, p.stdout[-2000:]
