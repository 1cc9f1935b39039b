// One-shot all-reduce of the per-step exchange over NVLink peer memory (one box, one process per GPU).
//
// What it replaces: the NCCL all-reduce of the packed float32 buffer [dU | dN | sum(R^2) | count] that ends every
// sharded Collocation_Loss step (pde_read_b200/parallel.py; reference semantics: the mean of Loss_Functions.py:65 and
// the parameter-gradient sums of Loss.backward(), Test_Train.py:99-100, over ALL points).  The payload is 84-330 KB:
// latency-bound, and NCCL's tree costs ~35 us for it on 8 GPUs.  Here every rank
//   push: writes its buffer straight into a slot of every peer's receive area (remote stores over NVLink), then
//         raises a flag in every peer;
//   pull: waits for the world's flags in its own area and sums the slots in rank order (bit-identical on all ranks).
// One kernel on the caller's stream, no host synchronisation, no NCCL.  Receive areas are double-buffered by the
// parity of an epoch counter that lives in device memory (so the launches can be captured in a CUDA graph): a rank can
// only be one step ahead of a peer, because finishing step e needs every peer's push of step e, which that peer
// enqueues after its own pull of step e-1.
//
// Layout of one rank's area (cudaMalloc'ed here so that its IPC handle covers exactly this allocation):
//   [0, 512)      header: epoch, done counter, flags[2][PEER_MAX_WORLD]
//   [512, ...)    float data[2 parities][world][cap]
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "internal.h"

namespace pdr {

constexpr int PEER_MAX_WORLD = 32;
constexpr int PEER_HDR_BYTES = 512;

struct PeerHeader {
    unsigned epoch;                          // last epoch this rank has pushed
    unsigned done;                           // blocks of the running push kernel that have finished their stores
    unsigned pad[62];
    unsigned flags[2][PEER_MAX_WORLD];       // flags[parity][r] = last epoch of that parity rank r has pushed to us
};
static_assert(sizeof(PeerHeader) == PEER_HDR_BYTES, "header layout");

__device__ __forceinline__ float* peer_slot(void* area, int parity, int world, int rank, int cap) {
    return reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(area) + PEER_HDR_BYTES) +
           ((size_t)parity * world + rank) * cap;
}

// One launch: push, publish, wait, sum.  Every block pushes its slice of the buffer to all peers; the last block to
// finish raises the flags; then every block waits for the world's flags in the local area and sums its slice.  The
// blocks spin, so the grid must be co-resident (<= 64 blocks of 256 threads).
__global__ void __launch_bounds__(256) peer_allreduce_kernel(float* __restrict__ buf, int n4, void* const* __restrict__ peers,
                                                            int rank, int world, int cap, void* local) {
    PeerHeader* h = reinterpret_cast<PeerHeader*>(local);
    __shared__ unsigned s_epoch;
    if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile unsigned*>(&h->epoch) + 1u;     // the epoch being produced
    __syncthreads();
    const unsigned e = s_epoch;
    const int par = (int)(e & 1u);
    const float4* s4 = reinterpret_cast<const float4*>(buf);
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = s4[i];
        for (int p = 0; p < world; ++p) {
            const int q = (rank + p) % world;            // spread the ranks' traffic over the links
            reinterpret_cast<float4*>(peer_slot(peers[q], par, world, rank, cap))[i] = v;
        }
    }
    __threadfence_system();                  // this thread's remote stores are visible system-wide ...
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&h->done, 1u);
        if (t == gridDim.x - 1) {            // ... and the last block to get here tells the peers
            h->done = 0u;
            __threadfence_system();
            for (int p = 0; p < world; ++p)
                *reinterpret_cast<volatile unsigned*>(&reinterpret_cast<PeerHeader*>(peers[p])->flags[par][rank]) = e;
        }
    }
    // (h->epoch is advanced by the block that sees the world's flags last -- see below -- so that no block of THIS
    //  launch can read the new value at its start)
    if (threadIdx.x < world) {
        // (bounded: ~10 s.  A peer that never arrives -- a crashed rank -- must not hang this GPU for ever; the sum is
        //  then garbage and h->pad[1] says so)
        const volatile unsigned* f = &h->flags[par][threadIdx.x];
        unsigned spins = 0;
        while (*f != e) {
            __nanosleep(64);
            if (++spins > (1u << 27)) { h->pad[1] = e; break; }
        }
    }
    __syncthreads();
    __threadfence();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < world; ++r) {    // fixed order: every rank gets the same bits
            const float4 v = __ldcg(reinterpret_cast<const float4*>(peer_slot(local, par, world, r, cap)) + i);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        reinterpret_cast<float4*>(buf)[i] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&h->pad[0], 1u);
        if (t == gridDim.x - 1) {
            h->pad[0] = 0u;
            h->epoch = e;                    // the next launch on this stream produces e + 1
        }
    }
}

}  // namespace pdr

using namespace pdr;

extern "C" {

size_t pderead_peer_area_bytes(int world, int cap_floats) {
    if (world < 1 || world > PEER_MAX_WORLD || cap_floats < 4) return 0;
    return (size_t)PEER_HDR_BYTES + sizeof(float) * 2 * (size_t)world * (size_t)((cap_floats + 3) & ~3);
}

int pderead_peer_alloc(int world, int cap_floats, void** d_area, unsigned char* handle64) {
    if (!d_area || !handle64) return PDEREAD_ERR_BAD_ARGUMENT;
    const size_t bytes = pderead_peer_area_bytes(world, cap_floats);
    if (!bytes) return PDEREAD_ERR_BAD_ARGUMENT;
    { const int rc_ = check_device(); if (rc_ != PDEREAD_OK) return rc_; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    PDR_CUDA_CHECK(cudaMalloc(&p, bytes));
    PDR_CUDA_CHECK(cudaMemset(p, 0, bytes));
    PDR_CUDA_CHECK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    PDR_CUDA_CHECK(cudaIpcGetMemHandle(&h, p));
    memcpy(handle64, &h, 64);
    *d_area = p;
    return PDEREAD_OK;
}

int pderead_peer_open(const unsigned char* handle64, void** d_area) {
    if (!handle64 || !d_area) return PDEREAD_ERR_BAD_ARGUMENT;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        // (a peer that is not visible to this process, or no peer access: the caller falls back to NCCL -- do not
        //  leave the error behind for the next launch check to trip over)
        set_last_cuda_error(e, __FILE__, __LINE__);
        (void)cudaGetLastError();
        return PDEREAD_ERR_CUDA;
    }
    *d_area = p;
    return PDEREAD_OK;
}

int pderead_peer_close(void* d_area) {
    if (!d_area) return PDEREAD_OK;
    PDR_CUDA_CHECK(cudaIpcCloseMemHandle(d_area));
    return PDEREAD_OK;
}

int pderead_peer_free(void* d_area) {
    if (!d_area) return PDEREAD_OK;
    PDR_CUDA_CHECK(cudaFree(d_area));
    return PDEREAD_OK;
}

int pderead_peer_allreduce(float* d_buf, int64_t n, void* const* d_peer_areas, void* d_local_area, int rank, int world,
                           int cap_floats, void* stream) {
    if (!d_buf || !d_peer_areas || !d_local_area || n < 0 || world < 1 || world > PEER_MAX_WORLD || rank < 0 ||
        rank >= world || (n & 3) || n > ((cap_floats + 3) & ~3) || (reinterpret_cast<uintptr_t>(d_buf) & 15))
        return PDEREAD_ERR_BAD_ARGUMENT;
    if (n == 0) return PDEREAD_OK;
    { const int rc_ = check_device(); if (rc_ != PDEREAD_OK) return rc_; }
    const int cap = (cap_floats + 3) & ~3;
    const int n4 = (int)(n >> 2);
    int blocks = (n4 + 255) / 256;
    if (blocks > 64) blocks = 64;            // co-resident by a wide margin (the pull kernel spins on flags)
    cudaStream_t s = (cudaStream_t)stream;
    peer_allreduce_kernel<<<blocks, 256, 0, s>>>(d_buf, n4, d_peer_areas, rank, world, cap, d_local_area);
    PDR_CUDA_CHECK(cudaGetLastError());
    stage_mark(s, ST_OTHER);
    return PDEREAD_OK;
}

}  // extern "C"
