"""World-size-2 tests of the point-sharding logic on CPU (gloo).  The arithmetic of each shard is
injected from the oracle; what is under test is pde_read_b200.parallel: slicing, the packed
all-reduce with the loss riding in the tail, the Gram all-reduce, and invariance of the result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pde_read_oracle as orc
from tests.helpers import load_golden, rel_err, split_state


def test_shard_bounds_cover_everything_once():
    from pde_read_b200.parallel import shard_bounds
    for M in (0, 1, 7, 8, 1000, 100_003):
        for W in (1, 2, 3, 8):
            spans = [shard_bounds(M, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == M
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _flat(gs, gp, tail=4):
    parts = [g.reshape(-1).float() for g in list(gs.values()) + list(gp.values())]
    return torch.cat(parts + [torch.zeros(tail)])


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pde_read_b200.parallel import ShardReducer, shard_points, sharded_loss_and_grads
        torch.set_num_threads(2)
        z = load_golden("random_tanh_m1n2.npz")
        sol, pde = split_state(z, "sol."), split_state(z, "pde.")
        P = torch.from_numpy(z["coords"])            # 96 points, identical on both ranks
        red = ShardReducer()
        assert red.world_size == world and red.rank == rank
        assert red.total_points(shard_points(P, rank, world).shape[0]) == P.shape[0]

        def local_fn(shard):
            loss, gs, gp = orc.collocation_loss_and_grads(sol, "Tanh", pde, "Tanh", 1, 2, shard, total_points=1)
            return torch.tensor([loss], dtype=torch.float64), _flat(gs, gp)

        loss, flat = sharded_loss_and_grads(local_fn, P, red)
        # Gram all-reduce
        b, Lib, _, _ = orc.generate_library(sol, "Tanh", pde, "Tanh", 1, 2, shard_points(P, rank, world), 2)
        A = torch.from_numpy(Lib).double()
        bb = torch.from_numpy(b).double()
        G, Atb, btb = red.reduce_gram(A.T @ A, A.T @ bb, (bb * bb).sum().reshape(1))
        q.put((rank, loss, flat[:-4].numpy(), G.numpy(), Atb.numpy(), float(btb)))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_step_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    z = load_golden("random_tanh_m1n2.npz")
    sol, pde = split_state(z, "sol."), split_state(z, "pde.")
    P = torch.from_numpy(z["coords"])
    loss1, gs, gp = orc.collocation_loss_and_grads(sol, "Tanh", pde, "Tanh", 1, 2, P)
    flat1 = _flat(gs, gp, tail=0).numpy()
    b, Lib, _, _ = orc.generate_library(sol, "Tanh", pde, "Tanh", 1, 2, P, 2)
    A = Lib.astype(np.float64)
    for rank, loss, flat, G, Atb, btb in results:
        assert abs(loss - loss1) <= 1e-6 * abs(loss1)
        assert abs(loss - float(z["loss"])) <= 1e-5 * float(z["loss"])
        assert rel_err(flat, flat1) < 2e-5
        assert rel_err(G, A.T @ A) < 1e-12 and rel_err(Atb, A.T @ b.astype(np.float64)) < 1e-12
        assert abs(btb - float((b.astype(np.float64) ** 2).sum())) < 1e-10
    # both ranks end with bit-identical reduced buffers (so they take identical optimizer steps)
    assert np.array_equal(results[0][2], results[1][2])


def _worker_totals(rank, world, port, q):
    """Two point sets whose totals differ (95 and 94) while rank 1's shard size is the same (47): the advisor's
    stale-total scenario.  Every rank must still issue the same collectives and divide by the right count."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pde_read_b200.parallel import ShardReducer, shard_points, sharded_loss_and_grads
        torch.set_num_threads(2)
        z = load_golden("random_tanh_m1n2.npz")
        sol, pde = split_state(z, "sol."), split_state(z, "pde.")
        P = torch.from_numpy(z["coords"])
        red = ShardReducer()

        def local_fn(shard):
            loss, gs, gp = orc.collocation_loss_and_grads(sol, "Tanh", pde, "Tanh", 1, 2, shard, total_points=1)
            return torch.tensor([loss], dtype=torch.float64), _flat(gs, gp)

        out = []
        for M in (95, 94, 95):
            sizes = shard_points(P[:M], rank, world).shape[0]
            loss, flat = sharded_loss_and_grads(local_fn, P[:M], red)
            out.append((M, sizes, loss, flat[:-4].numpy().copy(), red.total_points(sizes)))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_two_rank_totals_follow_the_data_not_a_cache():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_totals, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    z = load_golden("random_tanh_m1n2.npz")
    sol, pde = split_state(z, "sol."), split_state(z, "pde.")
    P = torch.from_numpy(z["coords"])
    assert [o[1] for o in results[1]] == [47, 47, 47]          # same shard size on rank 1 for both totals
    assert [o[1] for o in results[0]] == [48, 47, 48]
    for M_idx, M in enumerate((95, 94, 95)):
        loss1, gs, gp = orc.collocation_loss_and_grads(sol, "Tanh", pde, "Tanh", 1, 2, P[:M])
        flat1 = _flat(gs, gp, tail=0).numpy()
        for rank in (0, 1):
            m_, _, loss, flat, total = results[rank][M_idx]
            assert total == M
            assert abs(loss - loss1) <= 1e-6 * abs(loss1), (rank, M)
            assert rel_err(flat, flat1) < 2e-5, (rank, M)
