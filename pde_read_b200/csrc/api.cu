// C-ABI entry points of libpderead_b200.so for the collocation path (see include/pderead_b200.h).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>

#include "internal.h"
#include "umma.cuh"

namespace pdr {

static thread_local int g_last_cuda_error = 0;
static thread_local char g_last_cuda_where[160] = "";
void set_last_cuda_error(cudaError_t e, const char* file, int line) {
    g_last_cuda_error = (int)e;
    const char* base = file;
    for (const char* p = file; *p; ++p)
        if (*p == '/') base = p + 1;
    snprintf(g_last_cuda_where, sizeof(g_last_cuda_where), "%s:%d: %s", base, line, cudaGetErrorString(e));
    (void)cudaGetLastError();      // clear the non-sticky error so that the next call starts clean
}

// Optional measurement hook (bench.py): a caller-owned array of cudaEvent_t that the next calls on
// this host thread record, one after every kernel launch, so individual kernels can be timed
// inside an otherwise unmodified step without synchronising.
int ensure_dynamic_smem(const void* kernel, size_t smem) {
    struct Limit { const void* k; int dev; size_t smem; };
    constexpr int NLIMIT = 256;                  // > number of kernel instantiations x devices of one node in practice
    static Limit limits[NLIMIT];
    static int nlimits = 0;
    static std::mutex mu;
    int dev = 0;
    PDR_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    Limit* lim = nullptr;
    for (int i = 0; i < nlimits; ++i)
        if (limits[i].k == kernel && limits[i].dev == dev) {
            lim = &limits[i];
            break;
        }
    if (lim != nullptr && smem <= lim->smem) return PDEREAD_OK;
    if (lim == nullptr && nlimits == NLIMIT) {
        // table full: set the device maximum once for kernels that do not fit (never lowers anything)
        int optin = 0;
        PDR_CUDA_CHECK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        PDR_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        return PDEREAD_OK;
    }
    PDR_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (lim != nullptr) lim->smem = smem;
    else limits[nlimits++] = Limit{kernel, dev, smem};
    return PDEREAD_OK;
}

static thread_local cudaEvent_t* g_stage_events = nullptr;
static thread_local int g_stage_count = 0, g_stage_next = 0;
constexpr int MAX_STAGE_TAGS = 1024;
static thread_local unsigned char g_stage_tags[MAX_STAGE_TAGS];
void stage_mark(cudaStream_t s, int tag) {
    if (g_stage_events && g_stage_next < g_stage_count) {
        if (g_stage_next < MAX_STAGE_TAGS) g_stage_tags[g_stage_next] = (unsigned char)tag;
        cudaEventRecord(g_stage_events[g_stage_next++], s);
    }
}

// The kernels are built for sm_100a only: any other device gets PDEREAD_ERR_NOT_BLACKWELL instead of a launch
// failure.  One attribute query per device and process (remembered in an atomic table).
int check_device() {
    static std::atomic<int> state[64];      // 0 unknown, 1 ok, 2 not Blackwell
    int dev = 0;
    PDR_CUDA_CHECK(cudaGetDevice(&dev));
    const int slot = dev & 63;
    int st = state[slot].load(std::memory_order_relaxed);
    if (st == 0) {
        int major = 0;
        PDR_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
        st = (major == 10) ? 1 : 2;
        state[slot].store(st, std::memory_order_relaxed);
    }
    return st == 1 ? PDEREAD_OK : PDEREAD_ERR_NOT_BLACKWELL;
}

int block_warps(int W) {
    const int NC = (W + TN - 1) / TN;
    return NC < MAX_WARPS ? NC : MAX_WARPS;
}
// Reverse kernel: when the tile's two W x (J*TP) buffers leave room for a single CTA per SM, launch 16 warps
// -- the products use one warp per neuron chunk, the activation phases (latency-bound) use all of them.
int bwd_block_warps(int W, int L, int din, int J) {
    const int nw = block_warps(W);
    static int wide = -1;     // PDEREAD_BWD_WIDE=0: never the 16-warp block (A/B measurements)
    if (wide < 0) {
        const char* e = getenv("PDEREAD_BWD_WIDE");
        wide = (e && e[0] == '0') ? 0 : 1;
    }
    if (!wide) return nw;
    return (bwd_smem_bytes(W, L, din, J) > 112 * 1024 && nw < BWD_WIDE_WARPS) ? BWD_WIDE_WARPS : nw;
}
size_t fwd_smem_bytes_tile(int W, int din, int J, int nwarps, int TP) {
    const int JP = (use_pair_map(J) && J > 1) ? ((J + 1) & ~1) : J;    // ProdTile::JP_FWD
    const int HS = JP * TP + 4;
    return sizeof(float) * ((size_t)2 * W * HS + (size_t)din * TP + (size_t)nwarps * J * TP);
}
size_t fwd_smem_bytes(int W, int din, int J, int nwarps) { return fwd_smem_bytes_tile(W, din, J, nwarps, tile_points(J)); }
static const size_t SMEM_MAX = 227 * 1024;    // opt-in dynamic shared memory per CTA on sm_100
size_t fwd_hbuf_floats(int W, int J) {
    const int JP = (use_pair_map(J) && J > 1) ? ((J + 1) & ~1) : J;
    return (size_t)2 * W * (JP * tile_points(J) + 4);
}
size_t bwd_hbuf_floats(int W, int J) { return (size_t)2 * W * h_stride(J); }
size_t fwd_smem_small(int W, int din, int J, int nwarps) {
    return fwd_smem_bytes(W, din, J, nwarps) - sizeof(float) * fwd_hbuf_floats(W, J);
}
bool fwd_spills(int W, int din, int J) { return fwd_smem_bytes(W, din, J, block_warps(W)) > SMEM_MAX; }
static size_t bwd_smem_bytes_base(int W, int L, int din, int J) {
    const int TP = tile_points(J), HS = h_stride(J);
    const size_t ns = (((size_t)L * W + 1 + (size_t)W * din + W + 7 * L + 3) & ~(size_t)3) + HS;   // + the ones row
    return sizeof(float) * ((size_t)2 * W * HS + (size_t)2 * din * TP + (size_t)J * TP + ns);
}
static size_t bwd_wstage_bytes(int W, int J) {
    const int NC = (W + TN - 1) / TN;
    return sizeof(float) * (size_t)W * NC * (use_pair_map(J) ? TNP_PAIR : TNP_LANE);
}
// Forward kernel: per-warp weight slices in shared memory where they fit without costing a resident CTA (width 100,
// J = 6: one CTA per SM either way).
size_t fwd_wstage_bytes(int W, int L, int J, size_t base_smem) {
    static int mode = -1;     // PDEREAD_FWD_STAGE=0|1 overrides the rule below (A/B measurements)
    if (mode < 0) {
        const char* e = getenv("PDEREAD_FWD_STAGE");
        mode = !e ? 2 : (e[0] == '1' ? 1 : 0);
    }
    const int NC = (W + TN - 1) / TN;
    if (!mode || L < 2 || NC > MAX_WARPS) return 0;
    const size_t extra = bwd_wstage_bytes(W, J) + 16, sm = 227 * 1024;
    if (base_smem + extra > sm) return 0;
    return (mode == 1 || sm / (base_smem + extra + 1024) == sm / (base_smem + 1024)) ? extra : 0;
}
bool bwd_stage_weights(int W, int L, int din, int J) {
    static int mode = -1;     // PDEREAD_BWD_STAGE=0|1 overrides the rule below (A/B measurements)
    if (mode < 0) {
        const char* e = getenv("PDEREAD_BWD_STAGE");
        mode = !e ? 2 : (e[0] == '1' ? 1 : 0);
    }
    if (!mode || L < 2) return false;
    const size_t sm = 227 * 1024;
    const size_t base = bwd_smem_bytes_base(W, L, din, J) + 1024, staged = base + bwd_wstage_bytes(W, J) + 16;
    if (staged > sm) return false;
    // default: only where the copy does not cost a resident CTA (C4: -10 % on the reverse kernel; C2 / C3 would drop
    // from 4 / 3 CTAs per SM to 3 / 2 and lose 11 / 15 %)
    return mode == 1 || sm / staged == sm / base;
}
size_t bwd_smem_bytes(int W, int L, int din, int J) {
    return bwd_smem_bytes_base(W, L, din, J) + (bwd_stage_weights(W, L, din, J) ? bwd_wstage_bytes(W, J) + 16 : 0);
}
bool bwd_spills(int W, int L, int din, int J) { return bwd_smem_bytes(W, L, din, J) > SMEM_MAX; }
size_t bwd_smem_small(int W, int L, int din, int J) {
    return bwd_smem_bytes_base(W, L, din, J) - sizeof(float) * bwd_hbuf_floats(W, J);
}

// ---------------------------------------------------------------------------------------------
// small helper kernels
// ---------------------------------------------------------------------------------------------

// pair tile:  wt[l][k][c*TNP_PAIR+b*THP+n] = W_{l+1}[c*TN+b*TH+n][k]   (transposed, for z = W h)
//             wn[l][i][c*TNP_PAIR+b*THP+n] = W_{l+1}[i][c*TN+b*TH+n]   (native,     for hbar = W^T zbar)
// lane tile:  wt[l][k][c*TNP_LANE+n] = W_{l+1}[c*TN+n][k],  wn[l][i][c*TNP_LANE+n] = W_{l+1}[i][c*TN+n]
__global__ void pack_weights_kernel(NetDev net, int pair, float* wt, float* wn);

// ---- element functions of the two weight layouts (shared by the single-job kernels and pack_jobs_kernel) --------
__device__ __forceinline__ void pack_jet_element(const NetDev& net, int pair, float* wt, float* wn, size_t idx) {
    const int W = net.W, NC = net.NC;
    const int TNPk = pair ? TNP_PAIR : TNP_LANE;
    const size_t per_layer = (size_t)W * NC * TNPk;
    const int l = (int)(idx / per_layer);
    const size_t rem = idx - (size_t)l * per_layer;
    const int row = (int)(rem / (NC * TNPk));
    const int col = (int)(rem - (size_t)row * (NC * TNPk));
    const int c = col / TNPk, pos = col - c * TNPk;
    int o;
    bool ok;
    if (pair) {
        const int half = pos / THP, n = pos - half * THP;
        o = c * TN + half * TH + n;
        ok = (n < TH) && (o < W);
    } else {
        o = c * TN + pos;
        ok = (pos < TN) && (o < W);
    }
    const float* w = net.w[l + 1];
    if (wt) wt[idx] = ok ? w[(size_t)o * W + row] : 0.f;
    if (wn) wn[idx] = ok ? w[(size_t)row * W + o] : 0.f;
}
__device__ __forceinline__ void pack_plain_element(const NetDev& net, int WP, int native, float* wt, size_t idx) {
    const int W = net.W;
    const size_t per_layer = (size_t)W * WP;
    const int l = (int)(idx / per_layer);
    const size_t rem = idx - (size_t)l * per_layer;
    const int r = (int)(rem / WP), c = (int)(rem - (size_t)r * WP);
    float v = 0.f;
    if (c < W) v = native ? net.w[l + 1][(size_t)r * W + c] : net.w[l + 1][(size_t)c * W + r];
    wt[idx] = v;
}

__global__ void pack_weights_kernel(NetDev net, int pair, float* wt, float* wn) {
    const size_t total = (size_t)net.W * net.NC * (pair ? TNP_PAIR : TNP_LANE) * (net.L - 1);
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
        pack_jet_element(net, pair, wt, wn, idx);
}

// plain-MLP forward kernel (mlp_plain_kernels.cuh): wt[l][k][i] = W_{l+1}[i][k], rows padded with zeros to WP
// (native != 0: wn[l][i][k] = W_{l+1}[i][k], the layout of the plain reverse kernel)
__global__ void pack_plain_weights_kernel(NetDev net, int WP, int native, float* wt) {
    const size_t total = (size_t)net.W * WP * (net.L - 1);
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
        pack_plain_element(net, WP, native, wt, idx);
}

// All weight images of a fused step in ONE launch (blockIdx.y = job): the collocation step packs U (transposed +
// native, jet layout) and N (transposed + native, plain or jet layout) -- three or four ~10 us launches otherwise,
// 2 % of the C2 step and 5 % of the C1 step.
struct PackJob {
    NetDev net;
    int kind;        // 0: jet layout (a = wt, b = wn, arg = pair tile?), 1: plain layout (a = dst, arg = native?, wp)
    int arg, wp;
    float *a, *b;
    unsigned long long total;
};
struct PackJobs {
    PackJob j[4];
    int n;
};
__global__ void pack_jobs_kernel(const PackJobs jobs) {
    const PackJob& jb = jobs.j[blockIdx.y];
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < jb.total; idx += (size_t)gridDim.x * blockDim.x) {
        if (jb.kind == 0) pack_jet_element(jb.net, jb.arg, jb.a, jb.b, idx);
        else              pack_plain_element(jb.net, jb.wp, jb.arg, jb.a, idx);
    }
}

// dst = beta*dst + scale * sum_over_parts(gpart).  Block = 32 consecutive elements x 8 slices of the partial
// vectors; every thread sums its slice with four independent accumulators (the loads are the cost), the
// slices are combined through shared memory in a fixed order, so the result is bit-reproducible.
struct ReduceJob {
    const float* gpart;
    int nparts, stride;
    SegmentTable tab;
    float beta, scale;
};
__device__ __forceinline__ void reduce_partials_body(const ReduceJob& jb, float (&red)[8][33]) {
    const float* __restrict__ gpart = jb.gpart;
    const int nparts = jb.nparts, stride = jb.stride;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int e = blockIdx.x * 32 + tx;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (e < stride) {
        int p = ty;
        for (; p + 24 < nparts; p += 32) {
            a0 += gpart[(size_t)p * stride + e];
            a1 += gpart[(size_t)(p + 8) * stride + e];
            a2 += gpart[(size_t)(p + 16) * stride + e];
            a3 += gpart[(size_t)(p + 24) * stride + e];
        }
        for (; p < nparts; p += 8) a0 += gpart[(size_t)p * stride + e];
    }
    red[ty][tx] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (ty == 0 && e < stride) {
        int s = -1;
        for (int t = 0; t < jb.tab.n; ++t)
            if (e >= jb.tab.seg[t].off && e < jb.tab.seg[t].off + jb.tab.seg[t].count) { s = t; break; }
        if (s < 0) return;
        float acc = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) acc += red[y][tx];
        float* d = jb.tab.seg[s].dst + (e - jb.tab.seg[s].off);
        *d = (jb.beta == 0.f ? 0.f : jb.beta * (*d)) + jb.scale * acc;
    }
}
__global__ void __launch_bounds__(256) reduce_partials_kernel(const ReduceJob jb) {
    __shared__ float red[8][33];
    reduce_partials_body(jb, red);
}
// both networks of a fused step in one launch (blockIdx.y picks the network; same arithmetic, same fixed order)
__global__ void __launch_bounds__(256) reduce_partials2_kernel(const ReduceJob j0, const ReduceJob j1) {
    __shared__ float red[8][33];
    if (blockIdx.y == 0) {
        if (blockIdx.x * 32 < j0.stride) reduce_partials_body(j0, red);
    } else {
        if (blockIdx.x * 32 < j1.stride) reduce_partials_body(j1, red);
    }
}

// ---- weight images ------------------------------------------------------------------------------
// One image per product gi = 1..L of a network: [hi | lo], each MROWS x Kpad fp32, K-major core-matrix
// layout (8 rows x 16 bytes contiguous; LBO = 128 between 4-k groups, SBO = a_rs between 8-row groups).
// Row r of image gi < L holds the weights of neuron nid(r) of hidden layer gi (zero for unused rows);
// image L holds the output layer's weights in row 0.
__global__ void tc_pack_weights_kernel(NetDev net, int mrows, int Kpad, int NB, int a_rs, int a_bytes,
                                       float* __restrict__ img) {
    const int W = net.W, L = net.L;
    const size_t per_img = (size_t)2 * a_bytes / 4;   // floats
    const int per_mat = mrows * Kpad;
    const int nimg = L;   // L-1 hidden products + the output row
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (size_t)nimg * per_mat;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int gi = (int)(idx / per_mat) + 1;
        const int rem = (int)(idx - (size_t)(gi - 1) * per_mat);
        const int r = rem / Kpad, k = rem - r * Kpad;
        const int blk = (mrows == 64) ? 16 : 32;
        const int qd = r / blk, rl = r - qd * blk;
        const int nid = qd * NB + rl;
        float v = 0.f;
        if (k < W) {
            if (gi < L) {
                if (rl < NB && nid < W) v = net.w[gi][(size_t)nid * W + k];
            } else {
                if (r == 0) v = net.w[L][k];
            }
        }
        const size_t off = ((size_t)(r & 7) * 16 + (size_t)(r >> 3) * a_rs + (size_t)(k & 3) * 4 + (size_t)(k >> 2) * 128) / 4;
        float hi, lo;
        umma::tf32_split(v, hi, lo);
        float* base = img + (size_t)(gi - 1) * per_img;
        base[off] = hi;
        base[(size_t)a_bytes / 4 + off] = lo;
    }
}

// ---- loss tail of the sharded step: [sum(R^2) hi, lo | point count hi, lo] as float32 pairs that ride at the
// end of the packed gradient all-reduce (float32 SUM).  hi/lo of the sum keep ~48 bits; the count is split in
// 12-bit / high parts so that every partial sum stays an exact integer in float32.
__global__ void loss_tail_pack_kernel(const double* __restrict__ ss, long long M, float* __restrict__ tail) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double v = ss[0];
        const float hi = (float)v;
        tail[0] = hi;
        tail[1] = (float)(v - (double)hi);
        tail[2] = (float)(M >> 12);
        tail[3] = (float)(M & 4095);
    }
}
__global__ void loss_tail_unpack_kernel(const float* __restrict__ tail, double M_assumed, double* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double ssum = (double)tail[0] + (double)tail[1];
        const double cnt = (double)tail[2] * 4096.0 + (double)tail[3];
        out[0] = ssum;
        out[1] = cnt;
        out[2] = cnt > 0.0 ? M_assumed / cnt : 0.0;
    }
}

// ---- input normalisation of the PDE network ("PDE Network - Normalize Inputs", reference Network.py:100-104,
// 194-195: torch.nn.BatchNorm1d on the inputs of N) as a two-phase scheme --------------------------------------
// phase 1: moments of the inputs over ALL points (per-CTA float64 partial sums; the host all-reduces the 1 + 2 din
// numbers across ranks when the points are sharded); phase 2: the normalisation is an affine map of the inputs, so
// it is folded into the first Linear layer (W0' = W0 diag(s), b0' = b0 + W0 (beta - s mu), s = gamma rstd) and the
// unchanged MLP kernels run on the folded network.  The reverse pass maps the gradients of W0', b0' back to W0, b0,
// gamma, beta and adds the batch-statistics terms to the input gradient: gin_pd += a_d + b_d x_pd.
__global__ void __launch_bounds__(256) input_moments_kernel(const float* __restrict__ x, long long M, int din,
                                                            double* __restrict__ mom) {
    double s1[MAXD], s2[MAXD];
#pragma unroll
    for (int d = 0; d < MAXD; ++d) { s1[d] = 0.0; s2[d] = 0.0; }
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < M; p += (long long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int d = 0; d < MAXD; ++d)
            if (d < din) {
                const double v = (double)__ldg(x + p * din + d);
                s1[d] += v;
                s2[d] = fma(v, v, s2[d]);
            }
    }
    __shared__ double red[2 * MAXD][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 0; d < MAXD; ++d) {
        if (d < din) {
            double a = s1[d], b = s2[d];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                b += __shfl_xor_sync(0xffffffffu, b, o);
            }
            if (lane == 0) { red[d][warp] = a; red[MAXD + d][warp] = b; }
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * MAXD) {
        const int d = threadIdx.x % MAXD, which = threadIdx.x / MAXD;
        if (d < din) {
            double a = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += red[which * MAXD + d][w];
            atomicAdd(&mom[1 + which * din + d], a);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&mom[0], (double)M);
}

// stats = [mean (din) | rstd (din)]
__global__ void __launch_bounds__(256) bn_fold_kernel(NetDev net, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, const double* __restrict__ mom,
                                                      int training, float momentum, float eps,
                                                      float* __restrict__ running_mean, float* __restrict__ running_var,
                                                      long long* __restrict__ num_batches, float* __restrict__ w0f,
                                                      float* __restrict__ b0f, float* __restrict__ stats) {
    __shared__ float s_scale[MAXD], s_shift[MAXD];
    const int din = net.din, W = net.W;
    if (threadIdx.x < din) {
        const int d = threadIdx.x;
        float mean, var;
        if (training) {
            const double cnt = mom[0];
            const double mu = mom[1 + d] / cnt;
            double v = mom[1 + din + d] / cnt - mu * mu;        // biased variance normalises (torch BatchNorm1d)
            if (v < 0.0) v = 0.0;
            mean = (float)mu;
            var = (float)v;
            if (running_mean) running_mean[d] = (1.f - momentum) * running_mean[d] + momentum * mean;
            if (running_var) {
                const double unbiased = cnt > 1.0 ? v * cnt / (cnt - 1.0) : v;
                running_var[d] = (1.f - momentum) * running_var[d] + momentum * (float)unbiased;
            }
        } else {
            mean = running_mean[d];
            var = running_var[d];
        }
        const float rstd = rsqrtf(var + eps);
        stats[d] = mean;
        stats[din + d] = rstd;
        const float g = gamma ? gamma[d] : 1.f, b = beta ? beta[d] : 0.f;
        s_scale[d] = g * rstd;
        s_shift[d] = b - g * rstd * mean;
    }
    if (threadIdx.x == 0 && training && num_batches) num_batches[0] += 1;
    __syncthreads();
    for (int i = threadIdx.x; i < W; i += blockDim.x) {
        float acc = net.b[0][i];
        for (int d = 0; d < din; ++d) {
            const float w = net.w[0][(size_t)i * din + d];
            w0f[(size_t)i * din + d] = w * s_scale[d];
            acc = fmaf(w, s_shift[d], acc);
        }
        b0f[i] = acc;
    }
}

// sums[d] = sum_i gW0'[i][d] W0[i][d],  sums[din + d] = sum_i gb0'[i] W0[i][d]   (float64; one block)
__global__ void __launch_bounds__(256) bn_grad_sums_kernel(NetDev net, const float* __restrict__ gw0f,
                                                           const float* __restrict__ gb0f, double* __restrict__ sums) {
    const int din = net.din, W = net.W;
    __shared__ double red[8];
    for (int k = 0; k < 2 * din; ++k) {
        const int d = k % din, which = k / din;
        double a = 0.0;
        for (int i = threadIdx.x; i < W; i += blockDim.x) {
            const double w = (double)net.w[0][(size_t)i * din + d];
            a += w * (which == 0 ? (double)gw0f[(size_t)i * din + d] : (double)gb0f[i]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
            sums[k] = t;
        }
        __syncthreads();
    }
}

// In place on the gradient slots of Layers.0.weight (gW0' -> gW0); gb0 = gb0' needs no work.  gamma / beta gradients
// from the LOCAL sums (they are summed over ranks later with everything else); the input-gradient seeds a_d, b_d from
// the GLOBAL sums and the global count.
__global__ void __launch_bounds__(256) bn_backward_kernel(NetDev net, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const float* __restrict__ stats,
                                                          const double* __restrict__ mom,
                                                          const double* __restrict__ sums_local,
                                                          const double* __restrict__ sums_global, int training,
                                                          float* __restrict__ gw0, const float* __restrict__ gb0,
                                                          float* __restrict__ ggamma, float* __restrict__ gbeta,
                                                          float* __restrict__ seed_ab) {
    __shared__ float s_scale[MAXD], s_shift[MAXD];
    const int din = net.din, W = net.W;
    if (threadIdx.x < din) {
        const int d = threadIdx.x;
        const float mean = stats[d], rstd = stats[din + d];
        const float g = gamma ? gamma[d] : 1.f, b = beta ? beta[d] : 0.f;
        s_scale[d] = g * rstd;
        s_shift[d] = b - g * rstd * mean;
        {   // parameters of the normalisation layer: local sums
            const double S1 = sums_local[d], S2 = sums_local[din + d];
            const double gs = S1 - (double)mean * S2;           // d loss / d s_d at fixed statistics
            if (ggamma) ggamma[d] = (float)(gs * (double)rstd);
            if (gbeta) gbeta[d] = (float)S2;
        }
        float a = 0.f, bq = 0.f;
        if (training) {
            const double S1 = sums_global[d], S2 = sums_global[din + d];
            const double cnt = mom[0];
            const double gs = S1 - (double)mean * S2;
            const double gmu = -(double)(g * rstd) * S2;
            const double gvar = gs * (double)g * (-0.5 * (double)rstd * (double)rstd * (double)rstd);
            const double bd = 2.0 * gvar / cnt;
            bq = (float)bd;
            a = (float)(gmu / cnt - bd * (double)mean);
        }
        seed_ab[d] = a;
        seed_ab[din + d] = bq;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < W * din; idx += blockDim.x) {
        const int i = idx / din, d = idx - i * din;
        gw0[idx] = fmaf(gw0[idx], s_scale[d], gb0[i] * s_shift[d]);
    }
}

__global__ void __launch_bounds__(256) bn_input_grad_kernel(float* __restrict__ gin, const float* __restrict__ x,
                                                            long long total, int din, const float* __restrict__ seed_ab) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(idx % din);
        gin[idx] += fmaf(__ldg(seed_ab + din + d), __ldg(x + idx), __ldg(seed_ab + d));
    }
}

// ---- multi-tensor optimiser steps on one flat buffer (reference main.py:63-71: torch.optim.Adam / SGD(momentum 0.9,
// nesterov) over list(Sol_NN.parameters()) + list(PDE_NN.parameters()), stepped at Test_Train.py:105) -------------
// state[0] = step count (float, like torch's state['step']), state[1] = blocks finished (ticket).  Every block reads
// the old count, the last block to finish increments it: the kernel needs no host value that changes from step to
// step, so it can be captured in a CUDA graph and replayed.
// Same operation order as torch's single-tensor Adam: exp_avg.lerp_(g, 1-b1); exp_avg_sq.mul_(b2).addcmul_(g, g,
// 1-b2); denom = sqrt(exp_avg_sq) / sqrt(1-b2^t) + eps; p.addcdiv_(exp_avg, denom, -lr/(1-b1^t)).
__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, long long n,
                                                        float lr, float b1, float b2, float eps, float wd,
                                                        float grad_scale, float* __restrict__ state) {
    const float t = state[0] + 1.f;
    const double bc1 = 1.0 - pow((double)b1, (double)t), bc2 = 1.0 - pow((double)b2, (double)t);
    const float step_size = (float)((double)lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gi = g[i] * grad_scale;
        const float pi = p[i];
        if (wd != 0.f) gi = fmaf(wd, pi, gi);
        float mi = m[i], vi = v[i];
        mi = fmaf(gi - mi, 1.f - b1, mi);
        vi = fmaf(gi * gi, 1.f - b2, vi * b2);
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        m[i] = mi;
        v[i] = vi;
        p[i] = pi - step_size * (mi / denom);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int ticket = atomicAdd(reinterpret_cast<unsigned int*>(state + 1), 1u);
        if (ticket == gridDim.x - 1) {
            reinterpret_cast<unsigned int*>(state + 1)[0] = 0u;
            state[0] = t;
        }
    }
}

// torch.optim.SGD with momentum, dampening 0: buf = g on the first step, else buf = mom*buf + g;
// nesterov: d = g + mom*buf, else d = buf;  p -= lr*d.
__global__ void __launch_bounds__(256) sgd_step_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ buf, long long n, float lr, float mom,
                                                       int nesterov, float wd, float grad_scale,
                                                       float* __restrict__ state) {
    const bool first = (state[0] == 0.f);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gi = g[i] * grad_scale;
        const float pi = p[i];
        if (wd != 0.f) gi = fmaf(wd, pi, gi);
        float d = gi;
        if (mom != 0.f) {
            const float b = first ? gi : fmaf(mom, buf[i], gi);
            buf[i] = b;
            d = nesterov ? fmaf(mom, b, gi) : b;
        }
        p[i] = pi - lr * d;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int ticket = atomicAdd(reinterpret_cast<unsigned int*>(state + 1), 1u);
        if (ticket == gridDim.x - 1) {
            reinterpret_cast<unsigned int*>(state + 1)[0] = 0u;
            state[0] = state[0] + 1.f;
        }
    }
}

__global__ void fma_probe_kernel(int variant, int iters, float* sink) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f, c = 1e-3f;
    if (variant == 2) {
        // float64 FMA chains (roofline denominator of the Gram kernel)
        double d0 = a0, d1 = a1, d2 = a2, d3 = a3, d4 = a4, d5 = a5, d6 = a6, d7 = a7;
        const double dm = 0.999, dc = 1e-3;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                d0 = fma(d0, dm, dc); d1 = fma(d1, dm, dc); d2 = fma(d2, dm, dc); d3 = fma(d3, dm, dc);
                d4 = fma(d4, dm, dc); d5 = fma(d5, dm, dc); d6 = fma(d6, dm, dc); d7 = fma(d7, dm, dc);
            }
        }
        a0 = (float)(d0 + d1 + d2 + d3); a1 = (float)(d4 + d5 + d6 + d7);
    } else if (variant == 0) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
                a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
            }
        }
    } else {
        unsigned long long p0, p1, p2, p3, mm, cc;
        asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(a0), "f"(a1));
        asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(a2), "f"(a3));
        asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(a4), "f"(a5));
        asm("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(a6), "f"(a7));
        asm("mov.b64 %0, {%1, %2};" : "=l"(mm) : "f"(m), "f"(m));
        asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c), "f"(c));
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(mm), "l"(cc));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(mm), "l"(cc));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(mm), "l"(cc));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(mm), "l"(cc));
            }
        }
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(p0));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a2), "=f"(a3) : "l"(p1));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a4), "=f"(a5) : "l"(p2));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a6), "=f"(a7) : "l"(p3));
    }
    const float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123456.789f) sink[0] = s;   // keeps the chains alive; practically never true
}

// Philox-4x32-10 (Salmon et al. 2011): counter = (index, offset), key = seed
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t* out) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct BoundsDev { float lo[MAXD]; float span[MAXD]; };

__global__ void generate_points_kernel(BoundsDev bd, int dim, long long n, unsigned long long seed,
                                       unsigned long long offset, float* out) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n;
         p += (long long)gridDim.x * blockDim.x) {
        const unsigned long long ctr = offset + (unsigned long long)p;
        for (int d0 = 0; d0 < dim; d0 += 4) {
            uint32_t r[4];
            philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)(d0 >> 2), 0u, (uint32_t)seed,
                          (uint32_t)(seed >> 32), r);
            for (int t = 0; t < 4 && d0 + t < dim; ++t) {
                // 24 random bits -> [0,1); lo + u*span stays inside [lo, hi]
                const float u = (float)(r[t] >> 8) * (1.0f / 16777216.0f);
                out[p * dim + d0 + t] = fmaf(u, bd.span[d0 + t], bd.lo[d0 + t]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host-side helpers
// ---------------------------------------------------------------------------------------------
static int to_netdev(const pderead_net* n, NetDev* out) {
    if (n == nullptr) return PDEREAD_ERR_BAD_ARGUMENT;
    { const int rc_ = check_device(); if (rc_ != PDEREAD_OK) return rc_; }
    if (n->num_hidden_layers < 1 || n->num_hidden_layers > MAXL) return PDEREAD_ERR_UNSUPPORTED_NET;
    if (n->width < 1 || n->input_dim < 1 || n->input_dim > MAXD) return PDEREAD_ERR_UNSUPPORTED_NET;
    if (n->activation < 0 || n->activation > 2) return PDEREAD_ERR_UNSUPPORTED_NET;
    memset(out, 0, sizeof(*out));
    out->L = n->num_hidden_layers;
    out->W = n->width;
    out->din = n->input_dim;
    out->NC = (n->width + TN - 1) / TN;
    for (int i = 0; i <= out->L; ++i) {
        if (!n->weight[i] || !n->bias[i]) return PDEREAD_ERR_BAD_ARGUMENT;
        out->w[i] = n->weight[i];
        out->b[i] = n->bias[i];
    }
    for (int i = 0; i < out->L; ++i) {
        if (n->activation == PDEREAD_ACT_RATIONAL && (!n->rat_a[i] || !n->rat_b[i])) return PDEREAD_ERR_BAD_ARGUMENT;
        out->ra[i] = n->rat_a[i];
        out->rb[i] = n->rat_b[i];
    }
    return PDEREAD_OK;
}

static GradLayout make_layout(const NetDev& n) {
    GradLayout g;
    memset(&g, 0, sizeof(g));
    int off = 0;
    for (int i = 0; i <= n.L; ++i) {
        g.off_w[i] = off;
        const int in = (i == 0) ? n.din : n.W;
        const int outd = (i == n.L) ? 1 : n.W;
        off += in * outd;
        off = (off + 3) & ~3;
    }
    for (int i = 0; i <= n.L; ++i) {
        g.off_b[i] = off;
        off += (i == n.L) ? 1 : n.W;
    }
    off = (off + 3) & ~3;
    for (int i = 0; i < n.L; ++i) {
        g.off_ab[i] = off;
        off += 8;
    }
    g.total = (off + 3) & ~3;
    return g;
}

static int make_segments(const NetDev& n, int act, const GradLayout& g, const pderead_net_grad* gr, SegmentTable* t) {
    if (!gr) return PDEREAD_ERR_BAD_ARGUMENT;
    t->n = 0;
    for (int i = 0; i <= n.L; ++i) {
        const int in = (i == 0) ? n.din : n.W;
        const int outd = (i == n.L) ? 1 : n.W;
        if (!gr->weight[i] || !gr->bias[i]) return PDEREAD_ERR_BAD_ARGUMENT;
        t->seg[t->n++] = Segment{gr->weight[i], g.off_w[i], in * outd};
        t->seg[t->n++] = Segment{gr->bias[i], g.off_b[i], outd};
    }
    if (act == PDEREAD_ACT_RATIONAL) {
        for (int i = 0; i < n.L; ++i) {
            if (!gr->rat_a[i] || !gr->rat_b[i]) return PDEREAD_ERR_BAD_ARGUMENT;
            t->seg[t->n++] = Segment{gr->rat_a[i], g.off_ab[i], 4};
            t->seg[t->n++] = Segment{gr->rat_b[i], g.off_ab[i] + 4, 3};
        }
    }
    return PDEREAD_OK;
}

static bool order_supported(int nx, int nt) {
    if (nx < 0 || nx > PDEREAD_MAX_SPATIAL_ORDER || nt < 0 || nt > PDEREAD_MAX_TIME_ORDER) return false;
    if (nx == 0 && nt > 0) return false;
    return true;
}


// ---------------------------------------------------------------------------------------------
// tensor-core path: geometry, weight images, dispatch
// ---------------------------------------------------------------------------------------------
// Tensor-core policy (forward kernel only).  0 = off: CUDA-core kernels everywhere.  1 = auto (default): forward-only
// jet calls (Evaluate_Derivatives, PDE_Residual without grad, extraction) of networks wider than 64 with at least 3
// hidden layers run on the tensor cores (measured 1.5x faster at width 100, inside the stated derivative / residual
// tolerances).  Everything on the gradient path stays on the CUDA cores: the tensor core truncates inside its
// accumulation and 3xTF32 lands at ~1e-6 relative error per network instead of fp32's ~1e-7, which the residual
// cancellation of a trained network amplifies in the gradients.  2 = force: the tensor-core forward for
// every forward-only call whose shape it supports (tests, measurements).
static std::atomic<int> g_tc_mode{-1};
static int tc_mode() {
    int m = g_tc_mode.load(std::memory_order_relaxed);
    if (m < 0) {
        const char* e = getenv("PDEREAD_TC");
        m = 1;
        if (e && (e[0] == 'o' || e[0] == 'O' || e[0] == '0')) m = 0;
        if (e && (e[0] == 'f' || e[0] == 'F' || e[0] == '2')) m = 2;
        int expect = -1;
        if (!g_tc_mode.compare_exchange_strong(expect, m)) m = expect;
    }
    return m;
}
static bool tc_allowed(const NetDev& n, bool gradient_path) {
    const int m = tc_mode();
    if (m == 0 || gradient_path) return false;       // there is no tensor-core reverse kernel (removed in round 2)
    if (m == 2) return true;
    return n.W > 64 && n.L >= 3;   // (2 x 100 plain nets measured slower: the output product is half the work)
}

// Plain-MLP forward kernel (J = 1, register-tiled, weights staged in shared memory): on by default; PDEREAD_PLAIN_MLP=0
// or pderead_debug_set_plain_mlp(0) selects the lane = point kernel of mlp_jet_kernels.cuh instead (A/B measurements).
static std::atomic<int> g_plain_mode{-1};
static int plain_mode() {
    int m = g_plain_mode.load(std::memory_order_relaxed);
    if (m < 0) {
        const char* e = getenv("PDEREAD_PLAIN_MLP");
        m = (e && (e[0] == '0' || e[0] == 'o' || e[0] == 'O') && !(e[1] == 'n' || e[1] == 'N')) ? 0 : 1;
        int expect = -1;
        if (!g_plain_mode.compare_exchange_strong(expect, m)) m = expect;
    }
    return m;
}
static inline int plain_wp_h(int W) { return (W + 7) / 8 * 8; }
static bool plain_fits(const NetDev& n, bool stash) {
    if (!plain_mode() || (n.W + 7) / 8 > 16) return false;
    int dev = 0, maxsm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return false;
    const int PP = stash ? 2 : 4, TP = 32 * PP, HS = TP + 4, nw = (n.W + 7) / 8;
    const size_t smem = sizeof(float) * ((size_t)2 * n.W * HS + (size_t)n.din * TP + (size_t)nw * TP +
                                         (n.L > 1 ? (size_t)n.W * plain_wp_h(n.W) : 0));
    return smem + 1024 <= (size_t)maxsm;
}

static int round_up_i(int v, int a) { return (v + a - 1) / a * a; }

// Picks the tile geometry of the tensor-core kernels for one network and jet length; false if the
// shape is not supported (then the CUDA-core kernels run).
static bool tc_geometry(const NetDev& n, int J, bool gradient_path, TcGeom* g) {
    if (!tc_allowed(n, gradient_path)) return false;
    if (n.L < 2 || n.W < 8 || n.W > 128 || J < 1 || J > 7) return false;
    int dev = 0, maxsm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return false;
    memset(g, 0, sizeof(*g));
    g->mrows = (n.W <= 64) ? 64 : 128;
    const int subs = (g->mrows == 64) ? 2 : 1;
    g->Kpad = round_up_i(n.W, 8);
    g->NB = (n.W + 3) / 4;
    g->a_rs = (g->Kpad / 4) * 128;
    g->a_bytes = (g->mrows / 8) * g->a_rs;
    g->b_rs = (g->Kpad / 4) * 144;
    const int pcs_jet[] = {16, 8, 4, 0}, pcs_plain[] = {64, 32, 16, 8, 0};
    const int* pcs = (J == 1) ? pcs_plain : pcs_jet;
    for (int nwbuf = 2; nwbuf >= 1; --nwbuf) {
        for (int t = 0; pcs[t]; ++t) {
            const int PC = pcs[t];
            if ((PC / 2) % (J <= 4 ? 4 : 2) != 0) continue;   // the epilogue loads G points per TMEM access
            const int ncpad = round_up_i(PC * J, g->mrows == 128 ? 16 : 8);
            if (ncpad > 128) continue;
            const int slot = (ncpad / 8) * g->b_rs;
            const int sub_stride = 4 * slot + 64;
            const int TP = subs * 2 * PC;
            const size_t smem = (size_t)nwbuf * 2 * g->a_bytes + (size_t)subs * sub_stride +
                                (size_t)((n.din * TP + 1) & ~1) * 4 + 9 * 8 + 64;
            if (smem > (size_t)maxsm) continue;
            // prefer the larger chunk; with one weight buffer only if two do not fit at any chunk size >= 8
            if (nwbuf == 2 && PC < 8 && J > 1) continue;
            g->PC = PC;
            g->NCpad = ncpad;
            g->TP = TP;
            g->slot_bytes = slot;
            g->sub_stride = sub_stride;
            g->nwbuf = nwbuf;
            int cols = 32;
            while (cols < 4 * ncpad) cols <<= 1;     // 2 chunks x (main + correction accumulators)
            g->tmem_cols = cols;
            g->smem_bytes = smem;
            return true;
        }
    }
    return false;
}

// upper bound of the weight-image bytes of one network (L products, hi|lo, M = 128 rows)
static size_t tc_image_bytes_bound(const NetDev& n) {
    if (n.W > 128 || n.L < 2) return 0;
    return (size_t)n.L * 2 * 16 * (size_t)(round_up_i(n.W, 8) / 4) * 128;
}

static int tc_pack(const NetDev& n, int mrows, int Kpad, int NB, int a_rs, int a_bytes, float* img, cudaStream_t s) {
    const size_t total = (size_t)n.L * mrows * Kpad;
    const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    tc_pack_weights_kernel<<<blocks, 256, 0, s>>>(n, mrows, Kpad, NB, a_rs, a_bytes, img);
    PDR_CUDA_CHECK(cudaGetLastError());
    stage_mark(s, ST_PACK);
    return PDEREAD_OK;
}

#define PDR_FWD_COMBOS(X) X(0, 0) X(1, 0) X(2, 0) X(3, 0) X(4, 0) X(1, 1) X(2, 1) X(3, 1) X(4, 1) X(1, 2) X(2, 2) X(3, 2) X(4, 2)
#define PDR_BWD_COMBOS(X) X(0, 0) X(1, 1) X(2, 1) X(3, 1) X(4, 1) X(1, 2) X(2, 2) X(3, 2) X(4, 2)

static int dispatch_fwd_raw(int act, int nx, int nt, const FwdArgs& a, cudaStream_t s) {
#define CASE(X_, T_)                                                   \
    if (nx == X_ && nt == T_) {                                        \
        switch (act) {                                                 \
            case 0: return launch_fwd<0, X_, T_>(a, s);                \
            case 1: return launch_fwd<1, X_, T_>(a, s);                \
            case 2: return launch_fwd<2, X_, T_>(a, s);                \
        }                                                              \
    }
    PDR_FWD_COMBOS(CASE)
#undef CASE
    return PDEREAD_ERR_UNSUPPORTED_ORDER;
}

static int dispatch_bwd_grid(int act, int nx, int nt, int W, int L, int din, long long tiles, int* grid) {
#define CASE(X_, T_)                                                             \
    if (nx == X_ && nt == T_) {                                                  \
        switch (act) {                                                           \
            case 0: return bwd_grid<0, X_, T_>(W, L, din, tiles, grid);          \
            case 1: return bwd_grid<1, X_, T_>(W, L, din, tiles, grid);          \
            case 2: return bwd_grid<2, X_, T_>(W, L, din, tiles, grid);          \
        }                                                                        \
    }
    PDR_BWD_COMBOS(CASE)
#undef CASE
    return PDEREAD_ERR_UNSUPPORTED_ORDER;
}

static int dispatch_bwd_raw(int act, int nx, int nt, const BwdArgs& a, int grid, cudaStream_t s) {
#define CASE(X_, T_)                                                        \
    if (nx == X_ && nt == T_) {                                             \
        switch (act) {                                                      \
            case 0: return launch_bwd_impl<0, X_, T_>(a, grid, s);          \
            case 1: return launch_bwd_impl<1, X_, T_>(a, grid, s);          \
            case 2: return launch_bwd_impl<2, X_, T_>(a, grid, s);          \
        }                                                                   \
    }
    PDR_BWD_COMBOS(CASE)
#undef CASE
    return PDEREAD_ERR_UNSUPPORTED_ORDER;
}

static int dispatch_fwd(int act, int nx, int nt, const FwdArgs& a, cudaStream_t s) {
    const int rc = dispatch_fwd_raw(act, nx, nt, a, s);
    if (rc == PDEREAD_OK) stage_mark(s, (nx + nt) ? ST_FWD_JET : ST_FWD_PLAIN);
    return rc;
}
static int dispatch_bwd(int act, int nx, int nt, const BwdArgs& a, int grid, cudaStream_t s) {
    const int rc = dispatch_bwd_raw(act, nx, nt, a, grid, s);
    if (rc == PDEREAD_OK) stage_mark(s, (nx + nt) ? ST_BWD_JET : ST_BWD_PLAIN);
    return rc;
}

static int check_smem(int W, int L, int din, int J, bool bwd) {
    int dev = 0, maxsm = 0;
    PDR_CUDA_CHECK(cudaGetDevice(&dev));
    PDR_CUDA_CHECK(cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    size_t need = bwd ? bwd_smem_bytes(W, L, din, J) : fwd_smem_bytes(W, din, J, block_warps(W));
    if (need > (size_t)maxsm)     // activation buffers go to the workspace; the rest must still fit
        need = bwd ? bwd_smem_small(W, L, din, J) : fwd_smem_small(W, din, J, block_warps(W));
    return need <= (size_t)maxsm ? PDEREAD_OK : PDEREAD_ERR_UNSUPPORTED_NET;
}

// resident CTAs of the backward kernel for this network (upper bound on partial-gradient vectors)
static int max_grid(const NetDev& n, int J) {
    int dev = 0, sms = 148, maxsm = 232448;
    if (cudaGetDevice(&dev) == cudaSuccess) {
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    }
    if (bwd_spills(n.W, n.L, n.din, J)) return sms;
    const size_t smem = bwd_smem_bytes(n.W, n.L, n.din, J) + 1024;
    int occ = (int)((size_t)maxsm / smem);
    const int by_threads = 2048 / (block_warps(n.W) * 32);
    if (occ > by_threads) occ = by_threads;
    if (occ > 32) occ = 32;
    if (occ < 1) occ = 1;
    return sms * occ;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int sm_count() {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}
// workspace floats of the spilled activation buffers of one network (forward and, with need_backward, reverse kernel)
static size_t spill_floats_bound(const NetDev& n, int J, bool need_backward) {
    size_t f = 0;
    if (fwd_spills(n.W, n.din, J)) f += (size_t)sm_count() * fwd_hbuf_floats(n.W, J) + 64;
    if (need_backward && bwd_spills(n.W, n.L, n.din, J)) f += (size_t)sm_count() * bwd_hbuf_floats(n.W, J) + 64;
    return f;
}

// carve helper
struct Carver {
    char* base;
    size_t used, cap;
    bool ok;
    Carver(void* p, size_t c) : base((char*)p), used(0), cap(c), ok(true) {}
    template <typename T>
    T* take(size_t count) {
        const size_t bytes = align_up(count * sizeof(T), 256);
        if (used + bytes > cap) { ok = false; return nullptr; }
        T* r = (T*)(base + used);
        used += bytes;
        return r;
    }
};

static size_t packed_floats(const NetDev& n) {
    const int row = n.NC * TNP > plain_wp_h(n.W) ? n.NC * TNP : plain_wp_h(n.W);     // either weight layout fits
    return (size_t)(n.L > 1 ? n.L - 1 : 0) * n.W * row;
}
static size_t packed_floats_old(const NetDev& n) { return (size_t)(n.L > 1 ? n.L - 1 : 0) * n.W * n.NC * TNP; }
static size_t stash_floats_per_point(const NetDev& n, int J) { return (size_t)(n.L > 1 ? n.L - 1 : 0) * n.W * J; }
static size_t stash_floats_bound(const NetDev& n, int J) { return stash_floats_per_point(n, J); }

static int pack(const NetDev& n, int J, float* wt, float* wn, cudaStream_t s) {
    if (n.L <= 1 || (!wt && !wn)) return PDEREAD_OK;
    const size_t total = packed_floats_old(n);
    const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    pack_weights_kernel<<<blocks, 256, 0, s>>>(n, use_pair_map(J) ? 1 : 0, wt, wn);
    PDR_CUDA_CHECK(cudaGetLastError());
    stage_mark(s, ST_PACK);
    return PDEREAD_OK;
}

static int pack_plain(const NetDev& n, float* wt, cudaStream_t s, int native = 0) {
    if (n.L <= 1 || !wt) return PDEREAD_OK;
    const size_t total = (size_t)(n.L - 1) * n.W * plain_wp_h(n.W);
    const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    pack_plain_weights_kernel<<<blocks, 256, 0, s>>>(n, plain_wp_h(n.W), native, wt);
    PDR_CUDA_CHECK(cudaGetLastError());
    stage_mark(s, ST_PACK);
    return PDEREAD_OK;
}

static int dispatch_plain_fwd(int act, const FwdArgs& a, cudaStream_t s) {
    int rc = PDEREAD_ERR_UNSUPPORTED_NET;
    switch (act) {
        case 0: rc = launch_plain_fwd<0>(a, s); break;
        case 1: rc = launch_plain_fwd<1>(a, s); break;
        case 2: rc = launch_plain_fwd<2>(a, s); break;
    }
    if (rc == PDEREAD_OK) stage_mark(s, ST_FWD_PLAIN);
    return rc;
}

static bool plain_bwd_fits(const NetDev& n) {
    if (!plain_mode() || (n.W + 7) / 8 > 16) return false;
    int dev = 0, maxsm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return false;
    const int TP = 128, HS = TP + 4;
    const size_t ns = (((size_t)n.L * n.W + 1 + (size_t)n.W * n.din + n.W + 7 * n.L + 3) & ~(size_t)3) + HS;
    const size_t smem = sizeof(float) * ((size_t)2 * n.W * HS + (size_t)2 * n.din * TP + TP + ns +
                                         (n.L > 1 ? (size_t)n.W * plain_wp_h(n.W) : 0));
    return smem + 1024 <= (size_t)maxsm;
}

static int dispatch_plain_bwd_grid(int act, int W, int L, int din, long long tiles, int* grid) {
    switch (act) {
        case 0: return plain_bwd_grid<0>(W, L, din, tiles, grid);
        case 1: return plain_bwd_grid<1>(W, L, din, tiles, grid);
        case 2: return plain_bwd_grid<2>(W, L, din, tiles, grid);
    }
    return PDEREAD_ERR_UNSUPPORTED_NET;
}

static int dispatch_plain_bwd(int act, const BwdArgs& a, int grid, cudaStream_t s) {
    int rc = PDEREAD_ERR_UNSUPPORTED_NET;
    switch (act) {
        case 0: rc = launch_plain_bwd<0>(a, grid, s); break;
        case 1: rc = launch_plain_bwd<1>(a, grid, s); break;
        case 2: rc = launch_plain_bwd<2>(a, grid, s); break;
    }
    if (rc == PDEREAD_OK) stage_mark(s, ST_BWD_PLAIN);
    return rc;
}

static int reduce_partials(const float* gpart, int nparts, const GradLayout& lay, const SegmentTable& tab, float beta,
                           float scale, cudaStream_t s) {
    const int blocks = (lay.total + 31) / 32;
    ReduceJob jb;
    jb.gpart = gpart; jb.nparts = nparts; jb.stride = lay.total; jb.tab = tab; jb.beta = beta; jb.scale = scale;
    reduce_partials_kernel<<<blocks, 256, 0, s>>>(jb);
    PDR_CUDA_CHECK(cudaGetLastError());
    stage_mark(s, ST_REDUCE);
    return PDEREAD_OK;
}
static int reduce_partials2(const float* gp0, int np0, const GradLayout& l0, const SegmentTable& t0, const float* gp1,
                            int np1, const GradLayout& l1, const SegmentTable& t1, float beta, float scale, cudaStream_t s) {
    ReduceJob a, b;
    a.gpart = gp0; a.nparts = np0; a.stride = l0.total; a.tab = t0; a.beta = beta; a.scale = scale;
    b.gpart = gp1; b.nparts = np1; b.stride = l1.total; b.tab = t1; b.beta = beta; b.scale = scale;
    const int bx = ((l0.total > l1.total ? l0.total : l1.total) + 31) / 32;
    reduce_partials2_kernel<<<dim3(bx, 2), 256, 0, s>>>(a, b);
    PDR_CUDA_CHECK(cudaGetLastError());
    stage_mark(s, ST_REDUCE);
    return PDEREAD_OK;
}


static int dispatch_tc_fwd(int act, int nx, int nt, const TcFwdArgs& a, cudaStream_t s) {
#define CASE(X_, T_)                                                   \
    if (nx == X_ && nt == T_) {                                        \
        switch (act) {                                                 \
            case 0: { const int rc = launch_tc_fwd<0, X_, T_>(a, s); if (rc == PDEREAD_OK) stage_mark(s, (X_ + T_) ? ST_FWD_JET : ST_FWD_PLAIN); return rc; } \
            case 1: { const int rc = launch_tc_fwd<1, X_, T_>(a, s); if (rc == PDEREAD_OK) stage_mark(s, (X_ + T_) ? ST_FWD_JET : ST_FWD_PLAIN); return rc; } \
            case 2: { const int rc = launch_tc_fwd<2, X_, T_>(a, s); if (rc == PDEREAD_OK) stage_mark(s, (X_ + T_) ? ST_FWD_JET : ST_FWD_PLAIN); return rc; } \
        }                                                              \
    }
    PDR_FWD_COMBOS(CASE)
#undef CASE
    return PDEREAD_ERR_UNSUPPORTED_ORDER;
}

// A forward launch of one network on whichever path its plan selected.
struct FwdPlan {
    int J;        // jet length (selects the packed weight layout of the CUDA-core path)
    bool tc;
    bool plain;   // J = 1: register-tiled kernel with its own weight layout
    int stash_pp; // plain kernel with a stash: points per lane of the stash tile (2 or 4)
    TcGeom geom;
    float* wt;    // CUDA-core path: packed transposed weights
    float* img;   // tensor-core path: weight images
    float* hbuf;  // CUDA-core jet kernel: spilled activation buffers (networks too wide for shared memory), else null
};

static int plan_forward(const NetDev& n, int J, bool allow_tc, Carver& cv, FwdPlan* p) {
    memset(p, 0, sizeof(*p));
    p->J = J;
    // J = 1: the register-tiled plain-MLP kernel beats the tensor-core forward (5 x 100: 4.7 vs 6.0 ms per 2 M points)
    const bool prefer_plain = (J == 1) && tc_mode() != 2 && plain_fits(n, false);
    p->tc = allow_tc && !prefer_plain && tc_geometry(n, J, false, &p->geom);
    if (p->tc) {
        p->img = cv.take<float>((size_t)n.L * 2 * p->geom.a_bytes / 4);
    } else {
        p->plain = (J == 1) && plain_fits(n, false);
        p->wt = cv.take<float>(packed_floats(n));
        if (!p->plain && fwd_spills(n.W, n.din, J)) p->hbuf = cv.take<float>((size_t)sm_count() * fwd_hbuf_floats(n.W, J));
    }
    return cv.ok ? PDEREAD_OK : PDEREAD_ERR_WORKSPACE_TOO_SMALL;
}

static int plan_pack(const NetDev& n, const FwdPlan& p, cudaStream_t s) {
    if (p.tc) return tc_pack(n, p.geom.mrows, p.geom.Kpad, p.geom.NB, p.geom.a_rs, p.geom.a_bytes, p.img, s);
    if (p.plain) return pack_plain(n, p.wt, s);
    return pack(n, p.J, p.wt, nullptr, s);
}

// `f` carries everything except wt_packed / num_tiles, which depend on the path.
static int plan_launch(int act, int nx, int nt, const FwdPlan& p, FwdArgs f, cudaStream_t s) {
    if (!p.tc && p.plain) {
        f.wt_packed = p.wt;
        f.aux2 = p.stash_pp;
        return dispatch_plain_fwd(act, f, s);
    }
    if (!p.tc) {
        const int TP = tile_points(jet_len(nx, nt));
        f.wt_packed = p.wt;
        f.hbuf = p.hbuf;
        f.num_tiles = (f.M + TP - 1) / TP;
        return dispatch_fwd(act, nx, nt, f, s);
    }
    TcFwdArgs t;
    memset(&t, 0, sizeof(t));
    t.net = f.net;
    t.geom = p.geom;
    t.wimg = p.img;
    t.in = f.in;
    t.M = f.M;
    t.num_tiles = (f.M + p.geom.TP - 1) / p.geom.TP;
    t.out0 = f.out0;
    t.out1 = f.out1;
    t.dtm_in = f.dtm_in;
    t.resid_out = f.resid_out;
    t.seed_out = f.seed_out;
    t.grad_resid = f.grad_resid;
    t.seed_scale = f.seed_scale;
    t.sumsq = f.sumsq;
    return dispatch_tc_fwd(act, nx, nt, t, s);
}

constexpr long long CHUNK_ALIGN = 128;          // lcm of all tile sizes
constexpr long long CHUNK_TARGET = 131072;      // points per chunk the workspace query sizes for

static long long round_up_ll(long long v, long long a) { return (v + a - 1) / a * a; }

}  // namespace pdr

using namespace pdr;

extern "C" {

const char* pderead_version(void) { return "pderead_b200 0.1 (sm_100a)"; }

const char* pderead_status_string(int status) {
    switch (status) {
        case PDEREAD_OK: return "ok";
        case PDEREAD_ERR_BAD_ARGUMENT: return "bad argument";
        case PDEREAD_ERR_UNSUPPORTED_ORDER: return "derivative order not supported (space 0..4, time 0..2)";
        case PDEREAD_ERR_UNSUPPORTED_NET: return "network shape not supported (depth 1..16, input dim 1..8, width*jets must fit shared memory)";
        case PDEREAD_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
        case PDEREAD_ERR_CUDA: return "CUDA runtime error";
        case PDEREAD_ERR_NOT_BLACKWELL: return "device is not sm_100";
        default: return "unknown status";
    }
}

int pderead_last_cuda_error(void) { return g_last_cuda_error; }
const char* pderead_last_cuda_error_where(void) { return pdr::g_last_cuda_where; }

int pderead_debug_stage_events(void* const* events, int count) {
    g_stage_events = (cudaEvent_t*)events;
    g_stage_count = events ? count : 0;
    g_stage_next = 0;
    return PDEREAD_OK;
}
int pderead_debug_stage_events_recorded(void) { return g_stage_next; }
int pderead_debug_stage_tags(int32_t* h_tags, int capacity) {
    if (!h_tags || capacity < 0) return PDEREAD_ERR_BAD_ARGUMENT;
    int n = g_stage_next < MAX_STAGE_TAGS ? g_stage_next : MAX_STAGE_TAGS;
    if (n > capacity) n = capacity;
    for (int i = 0; i < n; ++i) h_tags[i] = g_stage_tags[i];
    return n;
}

int pderead_debug_set_plain_mlp(int on) {
    (void)plain_mode();
    return g_plain_mode.exchange(on ? 1 : 0);
}

int pderead_debug_set_tensor_cores(int mode) {
    (void)tc_mode();
    return g_tc_mode.exchange((mode < 0 || mode > 2) ? 1 : mode);
}

// ------------------------------------------------------------------------------------------
// Evaluate_Derivatives
// ------------------------------------------------------------------------------------------
size_t pderead_sol_derivatives_workspace_bytes(const pderead_net* sol, int m, int n, int64_t M, int need_backward) {
    NetDev U;
    if (to_netdev(sol, &U) != PDEREAD_OK || M < 0) return 0;
    const int J = jet_len(n, m);
    size_t bytes = align_up(packed_floats(U) * 4, 256) * (need_backward ? 2 : 1);
    bytes += align_up(spill_floats_bound(U, J, need_backward != 0) * 4, 256) + 512;
    bytes += align_up(tc_image_bytes_bound(U), 256);
    if (need_backward) {
        const GradLayout lay = make_layout(U);
        bytes += align_up((size_t)max_grid(U, J) * lay.total * 4, 256);
        const long long chunk = round_up_ll(M < CHUNK_TARGET ? (M > 0 ? M : 1) : CHUNK_TARGET, CHUNK_ALIGN);
        bytes += align_up((size_t)chunk * stash_floats_bound(U, J) * 4, 256);
        bytes += align_up((size_t)chunk * (n + 2) * 4, 256) * 2;
    }
    return bytes + 1024;
}

int pderead_sol_derivatives_forward(const pderead_net* sol, int m, int n, const float* d_coords, int64_t M,
                                    float* d_dtm, float* d_dxn, void* ws, size_t ws_bytes, void* stream) {
    NetDev U;
    int rc = to_netdev(sol, &U);
    if (rc) return rc;
    if (!order_supported(n, m)) return PDEREAD_ERR_UNSUPPORTED_ORDER;
    if (U.din != 2 && (n > 0 || m > 0)) return PDEREAD_ERR_UNSUPPORTED_NET;
    if (M < 0 || (M > 0 && (!d_coords || !d_dxn))) return PDEREAD_ERR_BAD_ARGUMENT;
    if (M == 0) return PDEREAD_OK;
    const int J = jet_len(n, m);
    cudaStream_t s = (cudaStream_t)stream;
    Carver cv(ws, ws_bytes);
    FwdPlan pl;
    if ((rc = plan_forward(U, J, true, cv, &pl))) return rc;
    if (!pl.tc && (rc = check_smem(U.W, U.L, U.din, J, false))) return rc;
    if (!ws && (pl.tc || packed_floats(U))) return PDEREAD_ERR_WORKSPACE_TOO_SMALL;
    if ((rc = plan_pack(U, pl, s))) return rc;
    FwdArgs a;
    memset(&a, 0, sizeof(a));
    a.net = U;
    a.in = d_coords;
    a.M = M;
    a.out0 = d_dxn;
    a.out1 = d_dtm;
    return plan_launch(sol->activation, n, m, pl, a, s);
}

// ---------------------------------------------------------------------------------------------
// reverse-pass plan of one network: tensor-core kernels (forward with stash + reverse) when the shape
// is supported, the CUDA-core kernels otherwise
// ---------------------------------------------------------------------------------------------
struct BwdPlan {
    bool plain;     // J = 1: register-tiled reverse kernel (tile of 128 points)
    FwdPlan fwd;
    float* wn;      // native packed weights
    float* hbuf;    // spilled Zb / Yb of the reverse kernel (networks too wide for shared memory), else null
    int grid;       // CTAs of the reverse kernel = number of partial-gradient vectors
    float* gpart;
    size_t stash_pp;   // stash floats per point
};

static int plan_backward(const NetDev& n, int act, int nx, int nt, bool want_gin, int64_t M, const GradLayout& lay,
                         Carver& cv, BwdPlan* p) {
    int rc;
    const int J = jet_len(nx, nt);
    memset(p, 0, sizeof(*p));
    p->fwd.J = J;
    (void)want_gin;
    {
        // (128-point tiles need enough points to fill the GPU: below ~100 tiles per 148 SMs the 64-point kernel wins,
        //  C1's 5000 points: 0.063 vs 0.081 ms)
        p->plain = (J == 1) && M >= 16384 && plain_bwd_fits(n) && plain_fits(n, false);
        if (!p->plain && (rc = check_smem(n.W, n.L, n.din, J, true))) return rc;
        p->fwd.plain = p->plain || ((J == 1) && plain_fits(n, true));
        p->fwd.stash_pp = p->plain ? 4 : 2;
        if (!p->fwd.plain && (rc = check_smem(n.W, n.L, n.din, J, false))) return rc;
        p->fwd.wt = cv.take<float>(packed_floats(n));
        p->wn = cv.take<float>(packed_floats(n));
        if (!p->fwd.plain && fwd_spills(n.W, n.din, J)) p->fwd.hbuf = cv.take<float>((size_t)sm_count() * fwd_hbuf_floats(n.W, J));
        if (!p->plain && bwd_spills(n.W, n.L, n.din, J)) p->hbuf = cv.take<float>((size_t)sm_count() * bwd_hbuf_floats(n.W, J));
        if (p->plain) {
            if ((rc = dispatch_plain_bwd_grid(act, n.W, n.L, n.din, (M + 127) / 128, &p->grid))) return rc;
        } else {
            const int TP = tile_points(J);
            if ((rc = dispatch_bwd_grid(act, nx, nt, n.W, n.L, n.din, (M + TP - 1) / TP, &p->grid))) return rc;
        }
        p->stash_pp = stash_floats_per_point(n, J);
    }
    p->gpart = cv.take<float>((size_t)p->grid * lay.total);
    return cv.ok ? PDEREAD_OK : PDEREAD_ERR_WORKSPACE_TOO_SMALL;
}

// ---- the same weight images as jobs of one pack_jobs_kernel launch (fused collocation step) -------------------
static void add_jet_job(PackJobs& js, const NetDev& n, int J, float* wt, float* wn) {
    if (n.L <= 1 || (!wt && !wn)) return;
    PackJob& j = js.j[js.n++];
    j.net = n; j.kind = 0; j.arg = use_pair_map(J) ? 1 : 0; j.wp = 0; j.a = wt; j.b = wn;
    j.total = (unsigned long long)(n.L - 1) * n.W * n.NC * (j.arg ? TNP_PAIR : TNP_LANE);
}
static void add_plain_job(PackJobs& js, const NetDev& n, float* dst, int native) {
    if (n.L <= 1 || !dst) return;
    PackJob& j = js.j[js.n++];
    j.net = n; j.kind = 1; j.arg = native; j.wp = plain_wp_h(n.W); j.a = dst; j.b = nullptr;
    j.total = (unsigned long long)(n.L - 1) * n.W * j.wp;
}
// false: this plan packs something the job kernel does not know (tensor-core weight images)
static bool forward_pack_jobs(const NetDev& n, const FwdPlan& p, PackJobs& js) {
    if (p.tc) return false;
    if (p.plain) add_plain_job(js, n, p.wt, 0);
    else         add_jet_job(js, n, p.J, p.wt, nullptr);
    return true;
}
static void backward_pack_jobs(const NetDev& n, const BwdPlan& p, PackJobs& js) {
    if (p.fwd.plain) {
        add_plain_job(js, n, p.fwd.wt, 0);
        if (p.plain) add_plain_job(js, n, p.wn, 1);
        else         add_jet_job(js, n, p.fwd.J, nullptr, p.wn);
    } else {
        add_jet_job(js, n, p.fwd.J, p.fwd.wt, p.wn);
    }
}
static int launch_pack_jobs(const PackJobs& js, cudaStream_t s) {
    if (js.n == 0) return PDEREAD_OK;
    unsigned long long most = 0;
    for (int i = 0; i < js.n; ++i) most = js.j[i].total > most ? js.j[i].total : most;
    const unsigned bx = (unsigned)((most + 255) / 256 < 592 ? (most + 255) / 256 : 592);
    pack_jobs_kernel<<<dim3(bx, js.n), 256, 0, s>>>(js);
    PDR_CUDA_CHECK(cudaGetLastError());
    stage_mark(s, ST_PACK);
    return PDEREAD_OK;
}

static int plan_backward_pack(const NetDev& n, const BwdPlan& p, cudaStream_t s) {
    if (p.fwd.plain) {
        int rc = pack_plain(n, p.fwd.wt, s);
        if (rc) return rc;
        return p.plain ? pack_plain(n, p.wn, s, 1) : pack(n, p.fwd.J, nullptr, p.wn, s);
    }
    return pack(n, p.fwd.J, p.fwd.wt, p.wn, s);
}

// `b` carries everything except the packed weights / tiles / grid, which depend on the path.
static int plan_backward_launch(int act, int nx, int nt, BwdPlan& p, BwdArgs b, bool first, cudaStream_t s) {
    b.accumulate = first ? 0 : 1;
    b.gpart = p.gpart;
    const int TP = p.plain ? 128 : tile_points(jet_len(nx, nt));
    b.wn_packed = p.wn;
    b.hbuf = p.plain ? nullptr : p.hbuf;
    b.num_tiles = (b.M + TP - 1) / TP;
    {
        // weight-gradient tiling: a warp owns 10 x (16 TK) outputs, TK in {2, 3, 4}; take the one with the fewest
        // (rounds x outputs per thread) over W x (W + 1 bias column); ties go to the larger tile (fewer
        // shared-memory wavefronts per FMA)
        const int W = b.net.W;
        const int nw = p.plain ? (W + 7) / 8 : bwd_block_warps(b.net.W, b.net.L, b.net.din, jet_len(nx, nt));
        long best = -1;
        for (int m = 2; m >= 0; --m) {
            const int tk = 2 + m;
            const int tasks = ((W + 9) / 10) * ((W + 1 + 16 * tk - 1) / (16 * tk));
            const long cost = (long)((tasks + nw - 1) / nw) * tk;
            if (best < 0 || cost < best) {
                best = cost;
                b.wg_mode = m;
            }
        }
    }
    int g = p.grid;
    if (b.num_tiles < g) g = (int)b.num_tiles;
    if (first) p.grid = g;
    if (p.plain) return dispatch_plain_bwd(act, b, g, s);
    return dispatch_bwd(act, nx, nt, b, g, s);
}

// shared implementation of "forward with stash + backward" over chunks for one network
static int net_backward_chunked(const NetDev& net, int act, int nx, int nt, const float* d_in, int64_t M,
                                const float* g0, const float* g1, float* d_gin, const pderead_net_grad* grads,
                                float beta, void* ws, size_t ws_bytes, cudaStream_t s) {
    int rc;
    const int J = jet_len(nx, nt);
    const GradLayout lay = make_layout(net);
    SegmentTable tab;
    if ((rc = make_segments(net, act, lay, grads, &tab))) return rc;

    Carver cv(ws, ws_bytes);
    BwdPlan pl;
    if ((rc = plan_backward(net, act, nx, nt, d_gin != nullptr, M, lay, cv, &pl))) return rc;
    // remaining bytes decide the chunk size
    const size_t per_point = (pl.stash_pp + (size_t)(nx + 2)) * 4;
    const size_t remain = ws_bytes - cv.used;
    long long chunk = (long long)((remain > 4096 ? remain - 4096 : 0) / (per_point ? per_point : 1));
    chunk = chunk / CHUNK_ALIGN * CHUNK_ALIGN;
    if (chunk < CHUNK_ALIGN) return PDEREAD_ERR_WORKSPACE_TOO_SMALL;
    if (chunk > round_up_ll(M, CHUNK_ALIGN)) chunk = round_up_ll(M, CHUNK_ALIGN);
    float* stash = cv.take<float>((size_t)chunk * pl.stash_pp);
    float* scratch0 = cv.take<float>((size_t)chunk * (nx + 1));
    float* scratch1 = cv.take<float>((size_t)chunk);
    if (!cv.ok) return PDEREAD_ERR_WORKSPACE_TOO_SMALL;
    if ((rc = plan_backward_pack(net, pl, s))) return rc;

    bool first = true;
    for (int64_t lo = 0; lo < M; lo += chunk) {
        const int64_t mc = (M - lo < chunk) ? (M - lo) : chunk;
        if (net.L > 1) {
            FwdArgs f;
            memset(&f, 0, sizeof(f));
            f.net = net;
            f.in = d_in + lo * net.din;
            f.M = mc;
            f.out0 = scratch0;
            f.out1 = scratch1;
            f.stash = stash;
            if ((rc = plan_launch(act, nx, nt, pl.fwd, f, s))) return rc;
        }
        BwdArgs b;
        memset(&b, 0, sizeof(b));
        b.net = net;
        b.lay = lay;
        b.in = d_in + lo * net.din;
        b.M = mc;
        b.stash = stash;
        b.g0 = g0 ? g0 + lo * (J == 1 ? 1 : (nx + 1)) : nullptr;
        b.g1 = g1 ? g1 + lo : nullptr;
        b.g0_scale = 1.f;
        b.g1_scale = 1.f;
        b.gin = d_gin ? d_gin + lo * net.din : nullptr;
        if ((rc = plan_backward_launch(act, nx, nt, pl, b, first, s))) return rc;
        first = false;
    }
    return reduce_partials(pl.gpart, pl.grid, lay, tab, beta, 1.f, s);
}

int pderead_sol_derivatives_backward(const pderead_net* sol, int m, int n, const float* d_coords, int64_t M,
                                     const float* d_grad_dtm, const float* d_grad_dxn, const pderead_net_grad* grads,
                                     float beta, void* ws, size_t ws_bytes, void* stream) {
    NetDev U;
    int rc = to_netdev(sol, &U);
    if (rc) return rc;
    if (!order_supported(n, m) || (n > 0 && m == 0)) return PDEREAD_ERR_UNSUPPORTED_ORDER;
    if (U.din != 2 && (n > 0 || m > 0)) return PDEREAD_ERR_UNSUPPORTED_NET;
    if (M <= 0 || !d_coords || !ws) return PDEREAD_ERR_BAD_ARGUMENT;
    return net_backward_chunked(U, sol->activation, n, m, d_coords, M, d_grad_dxn, d_grad_dtm, nullptr, grads, beta,
                                ws, ws_bytes, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------
// plain MLP (PDE_NN, and U without jets)
// ------------------------------------------------------------------------------------------
size_t pderead_mlp_workspace_bytes(const pderead_net* net, int64_t M, int need_backward) {
    NetDev N;
    if (to_netdev(net, &N) != PDEREAD_OK || M < 0) return 0;
    size_t bytes = align_up(packed_floats(N) * 4, 256) * (need_backward ? 2 : 1);
    bytes += align_up(spill_floats_bound(N, 1, need_backward != 0) * 4, 256) + 512;
    bytes += align_up(tc_image_bytes_bound(N), 256);
    if (need_backward) {
        const GradLayout lay = make_layout(N);
        bytes += align_up((size_t)max_grid(N, 1) * lay.total * 4, 256);
        const long long chunk = round_up_ll(M < CHUNK_TARGET ? (M > 0 ? M : 1) : CHUNK_TARGET, CHUNK_ALIGN);
        bytes += align_up((size_t)chunk * stash_floats_bound(N, 1) * 4, 256);
        bytes += align_up((size_t)chunk * 2 * 4, 256) * 2;
    }
    return bytes + 8192;
}

int pderead_mlp_forward(const pderead_net* net, const float* d_in, int64_t M, float* d_out, void* ws, size_t ws_bytes,
                        void* stream) {
    NetDev N;
    int rc = to_netdev(net, &N);
    if (rc) return rc;
    if (M < 0 || (M > 0 && (!d_in || !d_out))) return PDEREAD_ERR_BAD_ARGUMENT;
    if (M == 0) return PDEREAD_OK;
    cudaStream_t s = (cudaStream_t)stream;
    Carver cv(ws, ws_bytes);
    FwdPlan pl;
    if ((rc = plan_forward(N, 1, true, cv, &pl))) return rc;
    if (!pl.tc && !pl.plain && (rc = check_smem(N.W, N.L, N.din, 1, false))) return rc;
    if (!ws && (pl.tc || packed_floats(N))) return PDEREAD_ERR_WORKSPACE_TOO_SMALL;
    if ((rc = plan_pack(N, pl, s))) return rc;
    FwdArgs a;
    memset(&a, 0, sizeof(a));
    a.net = N;
    a.in = d_in;
    a.M = M;
    a.out0 = d_out;
    return plan_launch(net->activation, 0, 0, pl, a, s);
}

int pderead_mlp_backward(const pderead_net* net, const float* d_in, int64_t M, const float* d_grad_out,
                         float* d_grad_in, const pderead_net_grad* grads, float beta, void* ws, size_t ws_bytes,
                         void* stream) {
    NetDev N;
    int rc = to_netdev(net, &N);
    if (rc) return rc;
    if (M <= 0 || !d_in || !d_grad_out || !ws) return PDEREAD_ERR_BAD_ARGUMENT;
    return net_backward_chunked(N, net->activation, 0, 0, d_in, M, d_grad_out, nullptr, d_grad_in, grads, beta, ws,
                                ws_bytes, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------
// residual / collocation loss
// ------------------------------------------------------------------------------------------
size_t pderead_residual_workspace_bytes(const pderead_net* sol, const pderead_net* pde, int m, int n, int64_t M,
                                        int need_backward) {
    NetDev U, N;
    if (to_netdev(sol, &U) != PDEREAD_OK || to_netdev(pde, &N) != PDEREAD_OK || M < 0) return 0;
    const int J = jet_len(n, m);
    const long long chunk = round_up_ll(M < CHUNK_TARGET ? (M > 0 ? M : 1) : CHUNK_TARGET, CHUNK_ALIGN);
    size_t bytes = 0;
    bytes += align_up(packed_floats(U) * 4, 256) * (need_backward ? 2 : 1);
    bytes += align_up(packed_floats(N) * 4, 256) * (need_backward ? 2 : 1);
    bytes += align_up(spill_floats_bound(U, jet_len(n, m), need_backward != 0) * 4, 256) + 512;
    bytes += align_up(spill_floats_bound(N, 1, need_backward != 0) * 4, 256) + 512;
    bytes += align_up(tc_image_bytes_bound(U), 256) + align_up(tc_image_bytes_bound(N), 256);
    bytes += align_up((size_t)chunk * (n + 1) * 4, 256) + align_up((size_t)chunk * 4, 256);   // dxn, dtm
    if (need_backward) {
        const GradLayout lu = make_layout(U), ln = make_layout(N);
        bytes += align_up((size_t)max_grid(U, J) * lu.total * 4, 256) + align_up((size_t)max_grid(N, 1) * ln.total * 4, 256);
        bytes += align_up((size_t)chunk * stash_floats_bound(U, J) * 4, 256);
        bytes += align_up((size_t)chunk * stash_floats_bound(N, 1) * 4, 256);
        bytes += align_up((size_t)chunk * (n + 1) * 4, 256) + align_up((size_t)chunk * 4, 256);   // gdxn, seedN
    }
    return bytes + 8192;
}

static int residual_impl(const pderead_net* sol, const pderead_net* pde, int m, int n, const float* d_coords,
                         int64_t M, int64_t M_total, const float* d_grad_resid, float* d_resid, double* d_sum_sq,
                         const pderead_net_grad* sol_grads, const pderead_net_grad* pde_grads, float beta,
                         bool backward, void* ws, size_t ws_bytes, cudaStream_t s) {
    NetDev U, N;
    int rc;
    if ((rc = to_netdev(sol, &U))) return rc;
    if ((rc = to_netdev(pde, &N))) return rc;
    if (!order_supported(n, m) || n < 1 || m < 1) return PDEREAD_ERR_UNSUPPORTED_ORDER;
    if (U.din != 2 || N.din != n + 1) return PDEREAD_ERR_UNSUPPORTED_NET;
    if (M <= 0 || !d_coords || !ws || (M_total != 0 && M_total < M)) return PDEREAD_ERR_BAD_ARGUMENT;
    const int J = jet_len(n, m);
    const GradLayout lu = make_layout(U), ln = make_layout(N);
    SegmentTable tu, tn;
    if (backward) {
        if ((rc = make_segments(U, sol->activation, lu, sol_grads, &tu))) return rc;
        if ((rc = make_segments(N, pde->activation, ln, pde_grads, &tn))) return rc;
    }

    Carver cv(ws, ws_bytes);
    BwdPlan bU, bN;           // backward: forward-with-stash + reverse plans
    FwdPlan fU, fN;           // forward only
    memset(&bU, 0, sizeof(bU));
    memset(&bN, 0, sizeof(bN));
    if (backward) {
        if ((rc = plan_backward(U, sol->activation, n, m, false, M, lu, cv, &bU))) return rc;
        if ((rc = plan_backward(N, pde->activation, 0, 0, true, M, ln, cv, &bN))) return rc;
        fU = bU.fwd;
        fN = bN.fwd;
    } else {
        if ((rc = plan_forward(U, J, true, cv, &fU))) return rc;
        if ((rc = plan_forward(N, 1, true, cv, &fN))) return rc;
        if (!fU.tc && (rc = check_smem(U.W, U.L, U.din, J, false))) return rc;
        if (!fN.tc && !fN.plain && (rc = check_smem(N.W, N.L, N.din, 1, false))) return rc;
    }
    if (!cv.ok) return PDEREAD_ERR_WORKSPACE_TOO_SMALL;
    size_t per_point = (size_t)(n + 2) * 4;
    if (backward) per_point += (bU.stash_pp + bN.stash_pp + (size_t)(n + 2)) * 4;
    const size_t remain = ws_bytes - cv.used;
    long long chunk = (long long)((remain > 8192 ? remain - 8192 : 0) / per_point);
    chunk = chunk / CHUNK_ALIGN * CHUNK_ALIGN;
    if (chunk < CHUNK_ALIGN) return PDEREAD_ERR_WORKSPACE_TOO_SMALL;
    if (chunk > round_up_ll(M, CHUNK_ALIGN)) chunk = round_up_ll(M, CHUNK_ALIGN);
    float* dxn = cv.take<float>((size_t)chunk * (n + 1));
    float* dtm = cv.take<float>((size_t)chunk);
    float *stU = nullptr, *stN = nullptr, *gdxn = nullptr, *seedN = nullptr;
    if (backward) {
        stU = cv.take<float>((size_t)chunk * bU.stash_pp);
        stN = cv.take<float>((size_t)chunk * bN.stash_pp);
        gdxn = cv.take<float>((size_t)chunk * (n + 1));
        seedN = cv.take<float>((size_t)chunk);
    }
    if (!cv.ok) return PDEREAD_ERR_WORKSPACE_TOO_SMALL;

    {
        // one launch for every weight image of the step (the tensor-core forward's images keep their own kernel)
        PackJobs js;
        memset(&js, 0, sizeof(js));
        if (backward) {
            backward_pack_jobs(U, bU, js);
            backward_pack_jobs(N, bN, js);
        } else {
            if (!forward_pack_jobs(U, fU, js) && (rc = plan_pack(U, fU, s))) return rc;
            if (!forward_pack_jobs(N, fN, js) && (rc = plan_pack(N, fN, s))) return rc;
        }
        if ((rc = launch_pack_jobs(js, s))) return rc;
    }

    bool first = true;
    for (int64_t lo = 0; lo < M; lo += chunk) {
        const int64_t mc = (M - lo < chunk) ? (M - lo) : chunk;
        // U jets
        FwdArgs fu;
        memset(&fu, 0, sizeof(fu));
        fu.net = U;
        fu.in = d_coords + lo * 2;
        fu.M = mc;
        fu.out0 = dxn;
        fu.out1 = dtm;
        fu.stash = (backward && U.L > 1) ? stU : nullptr;
        if ((rc = plan_launch(sol->activation, n, m, fU, fu, s))) return rc;
        // N forward + residual epilogue
        FwdArgs fn;
        memset(&fn, 0, sizeof(fn));
        fn.net = N;
        fn.in = dxn;
        fn.M = mc;
        fn.stash = (backward && N.L > 1) ? stN : nullptr;
        fn.dtm_in = dtm;
        fn.resid_out = d_resid ? d_resid + lo : nullptr;
        fn.seed_out = backward ? seedN : nullptr;
        fn.grad_resid = d_grad_resid ? d_grad_resid + lo : nullptr;
        fn.seed_scale = M_total > 0 ? (float)(2.0 / (double)M_total) : 2.f;   // M_total = 0: unscaled sum(R^2)
        fn.sumsq = d_sum_sq;
        if ((rc = plan_launch(pde->activation, 0, 0, fN, fn, s))) return rc;
        if (backward) {
            BwdArgs bn;
            memset(&bn, 0, sizeof(bn));
            bn.net = N;
            bn.lay = ln;
            bn.in = dxn;
            bn.M = mc;
            bn.stash = stN;
            bn.g0 = seedN;
            bn.g0_scale = 1.f;
            bn.gin = gdxn;
            if ((rc = plan_backward_launch(pde->activation, 0, 0, bN, bn, first, s))) return rc;

            BwdArgs bu;
            memset(&bu, 0, sizeof(bu));
            bu.net = U;
            bu.lay = lu;
            bu.in = d_coords + lo * 2;
            bu.M = mc;
            bu.stash = stU;
            bu.g0 = gdxn;
            bu.g0_scale = 1.f;
            bu.g1 = seedN;          // dL/dDtm = dL/dR = -dL/dN
            bu.g1_scale = -1.f;
            if ((rc = plan_backward_launch(sol->activation, n, m, bU, bu, first, s))) return rc;
        }
        first = false;
    }
    if (backward) {
        if ((rc = reduce_partials2(bU.gpart, bU.grid, lu, tu, bN.gpart, bN.grid, ln, tn, beta, 1.f, s))) return rc;
    }
    return PDEREAD_OK;
}

int pderead_residual_forward(const pderead_net* sol, const pderead_net* pde, int m, int n, const float* d_coords,
                             int64_t M, float* d_resid, double* d_sum_sq, void* ws, size_t ws_bytes, void* stream) {
    return residual_impl(sol, pde, m, n, d_coords, M, M, nullptr, d_resid, d_sum_sq, nullptr, nullptr, 0.f, false, ws,
                         ws_bytes, (cudaStream_t)stream);
}

int pderead_collocation_loss_grad(const pderead_net* sol, const pderead_net* pde, int m, int n, const float* d_coords,
                                  int64_t M, int64_t M_total, const float* d_grad_resid, float* d_resid,
                                  double* d_sum_sq, const pderead_net_grad* sol_grads, const pderead_net_grad* pde_grads,
                                  float beta, void* ws, size_t ws_bytes, void* stream) {
    return residual_impl(sol, pde, m, n, d_coords, M, M_total, d_grad_resid, d_resid, d_sum_sq, sol_grads, pde_grads,
                         beta, true, ws, ws_bytes, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------
// extraction: coords -> Gram (reference Extraction.py:305-372 without materialising the library)
// ------------------------------------------------------------------------------------------
size_t pderead_extraction_gram_workspace_bytes(const pderead_net* sol, const pderead_net* pde, int n, int K,
                                               int64_t M) {
    NetDev U, N;
    (void)K;
    if (to_netdev(sol, &U) != PDEREAD_OK || to_netdev(pde, &N) != PDEREAD_OK || M < 0) return 0;
    const long long chunk = round_up_ll(M < 8 * CHUNK_TARGET ? (M > 0 ? M : 1) : 8 * CHUNK_TARGET, CHUNK_ALIGN);
    size_t bytes = align_up(packed_floats(U) * 4, 256) + align_up(packed_floats(N) * 4, 256);
    bytes += align_up(spill_floats_bound(U, jet_len(n, 0), false) * 4, 256) + align_up(spill_floats_bound(N, 1, false) * 4, 256) + 1024;
    bytes += align_up(tc_image_bytes_bound(U), 256) + align_up(tc_image_bytes_bound(N), 256);
    bytes += align_up((size_t)chunk * (n + 1) * 4, 256) + align_up((size_t)chunk * 4, 256);
    return bytes + 8192;
}

int pderead_extraction_gram(const pderead_net* sol, const pderead_net* pde, int n, int K, const float* d_coords,
                            int64_t M, double* d_G, double* d_Atb, double* d_btb, void* ws, size_t ws_bytes,
                            void* stream) {
    NetDev U, N;
    int rc;
    if ((rc = to_netdev(sol, &U))) return rc;
    if ((rc = to_netdev(pde, &N))) return rc;
    if (!order_supported(n, 0) || n < 1) return PDEREAD_ERR_UNSUPPORTED_ORDER;
    if (U.din != 2 || N.din != n + 1) return PDEREAD_ERR_UNSUPPORTED_NET;
    if (M <= 0 || !d_coords || !ws || !d_G || !d_Atb || !d_btb) return PDEREAD_ERR_BAD_ARGUMENT;
    if (pderead_library_num_columns(n, K) < 1) return PDEREAD_ERR_BAD_ARGUMENT;
    const int J = jet_len(n, 0);
    cudaStream_t s = (cudaStream_t)stream;
    Carver cv(ws, ws_bytes);
    FwdPlan plU, plN;
    if ((rc = plan_forward(U, J, true, cv, &plU))) return rc;
    if ((rc = plan_forward(N, 1, true, cv, &plN))) return rc;
    if (!plU.tc && (rc = check_smem(U.W, U.L, U.din, J, false))) return rc;
    if (!plN.tc && !plN.plain && (rc = check_smem(N.W, N.L, N.din, 1, false))) return rc;
    const size_t per_point = (size_t)(n + 2) * 4;
    const size_t remain = ws_bytes - cv.used;
    long long chunk = (long long)((remain > 4096 ? remain - 4096 : 0) / per_point);
    chunk = chunk / CHUNK_ALIGN * CHUNK_ALIGN;
    if (chunk < CHUNK_ALIGN) return PDEREAD_ERR_WORKSPACE_TOO_SMALL;
    if (chunk > round_up_ll(M, CHUNK_ALIGN)) chunk = round_up_ll(M, CHUNK_ALIGN);
    float* dxn = cv.take<float>((size_t)chunk * (n + 1));
    float* bb = cv.take<float>((size_t)chunk);
    if (!cv.ok) return PDEREAD_ERR_WORKSPACE_TOO_SMALL;
    if ((rc = plan_pack(U, plU, s))) return rc;
    if ((rc = plan_pack(N, plN, s))) return rc;
    for (int64_t lo = 0; lo < M; lo += chunk) {
        const int64_t mc = (M - lo < chunk) ? (M - lo) : chunk;
        FwdArgs fu;
        memset(&fu, 0, sizeof(fu));
        fu.net = U;
        fu.in = d_coords + lo * 2;
        fu.M = mc;
        fu.out0 = dxn;
        if ((rc = plan_launch(sol->activation, n, 0, plU, fu, s))) return rc;
        FwdArgs fn;
        memset(&fn, 0, sizeof(fn));
        fn.net = N;
        fn.in = dxn;
        fn.M = mc;
        fn.out0 = bb;
        if ((rc = plan_launch(pde->activation, 0, 0, plN, fn, s))) return rc;
        if ((rc = library_gram_launch(dxn, bb, mc, n, K, d_G, d_Atb, d_btb, s))) return rc;
    }
    return PDEREAD_OK;
}

// ------------------------------------------------------------------------------------------
// input normalisation of the PDE network, two-phase (see the kernels above)
// ------------------------------------------------------------------------------------------
int pderead_input_moments(const float* d_x, int64_t M, int din, double* d_moments, void* stream) {
    if (!d_moments || M < 0 || din < 1 || din > MAXD || (M > 0 && !d_x)) return PDEREAD_ERR_BAD_ARGUMENT;
    if (M == 0) return PDEREAD_OK;
    { const int rc_ = check_device(); if (rc_ != PDEREAD_OK) return rc_; }
    const long long blocks = (M + 255) / 256;
    input_moments_kernel<<<(int)(blocks < 592 ? blocks : 592), 256, 0, (cudaStream_t)stream>>>(d_x, (long long)M, din,
                                                                                                d_moments);
    PDR_CUDA_CHECK(cudaGetLastError());
    stage_mark((cudaStream_t)stream, ST_BN);
    return PDEREAD_OK;
}

int pderead_bn_fold(const pderead_net* net, const float* d_gamma, const float* d_beta, const double* d_moments,
                    int training, float momentum, float eps, float* d_running_mean, float* d_running_var,
                    int64_t* d_num_batches, float* d_w0_folded, float* d_b0_folded, float* d_stats, void* stream) {
    NetDev N;
    int rc = to_netdev(net, &N);
    if (rc) return rc;
    if (!d_w0_folded || !d_b0_folded || !d_stats) return PDEREAD_ERR_BAD_ARGUMENT;
    if (training ? !d_moments : (!d_running_mean || !d_running_var)) return PDEREAD_ERR_BAD_ARGUMENT;
    bn_fold_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(N, d_gamma, d_beta, d_moments, training, momentum, eps,
                                                        d_running_mean, d_running_var, (long long*)d_num_batches,
                                                        d_w0_folded, d_b0_folded, d_stats);
    PDR_CUDA_CHECK(cudaGetLastError());
    stage_mark((cudaStream_t)stream, ST_BN);
    return PDEREAD_OK;
}

int pderead_bn_grad_sums(const pderead_net* net, const float* d_gw0_folded, const float* d_gb0_folded, double* d_sums,
                         void* stream) {
    NetDev N;
    int rc = to_netdev(net, &N);
    if (rc) return rc;
    if (!d_gw0_folded || !d_gb0_folded || !d_sums) return PDEREAD_ERR_BAD_ARGUMENT;
    bn_grad_sums_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(N, d_gw0_folded, d_gb0_folded, d_sums);
    PDR_CUDA_CHECK(cudaGetLastError());
    stage_mark((cudaStream_t)stream, ST_BN);
    return PDEREAD_OK;
}

int pderead_bn_backward(const pderead_net* net, const float* d_gamma, const float* d_beta, const float* d_stats,
                        const double* d_moments, const double* d_sums_local, const double* d_sums_global, int training,
                        float* d_gw0, const float* d_gb0, float* d_ggamma, float* d_gbeta, float* d_seed_ab,
                        void* stream) {
    NetDev N;
    int rc = to_netdev(net, &N);
    if (rc) return rc;
    if (!d_stats || !d_sums_local || !d_sums_global || !d_gw0 || !d_gb0 || !d_seed_ab || (training && !d_moments))
        return PDEREAD_ERR_BAD_ARGUMENT;
    bn_backward_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(N, d_gamma, d_beta, d_stats, d_moments, d_sums_local,
                                                            d_sums_global, training, d_gw0, d_gb0, d_ggamma, d_gbeta,
                                                            d_seed_ab);
    PDR_CUDA_CHECK(cudaGetLastError());
    stage_mark((cudaStream_t)stream, ST_BN);
    return PDEREAD_OK;
}

int pderead_bn_input_grad(float* d_gin, const float* d_x, int64_t M, int din, const float* d_seed_ab, void* stream) {
    if (M < 0 || din < 1 || din > MAXD || !d_seed_ab || (M > 0 && (!d_gin || !d_x))) return PDEREAD_ERR_BAD_ARGUMENT;
    if (M == 0) return PDEREAD_OK;
    { const int rc_ = check_device(); if (rc_ != PDEREAD_OK) return rc_; }
    const long long total = (long long)M * din, blocks = (total + 255) / 256;
    bn_input_grad_kernel<<<(int)(blocks < 1184 ? blocks : 1184), 256, 0, (cudaStream_t)stream>>>(d_gin, d_x, total, din,
                                                                                                  d_seed_ab);
    PDR_CUDA_CHECK(cudaGetLastError());
    stage_mark((cudaStream_t)stream, ST_BN);
    return PDEREAD_OK;
}

// ------------------------------------------------------------------------------------------
// optimiser step over all parameters of both networks in one launch
// ------------------------------------------------------------------------------------------
static int optim_blocks(int64_t n) {
    const int64_t b = (n + 255) / 256;
    return (int)(b < 1 ? 1 : (b > 592 ? 592 : b));
}

int pderead_adam_step(float* d_params, const float* d_grads, float* d_exp_avg, float* d_exp_avg_sq, int64_t n, float lr,
                      float beta1, float beta2, float eps, float weight_decay, float grad_scale, float* d_state2,
                      void* stream) {
    if (!d_params || !d_grads || !d_exp_avg || !d_exp_avg_sq || !d_state2 || n < 0) return PDEREAD_ERR_BAD_ARGUMENT;
    if (n == 0) return PDEREAD_OK;
    { const int rc_ = check_device(); if (rc_ != PDEREAD_OK) return rc_; }
    adam_step_kernel<<<optim_blocks(n), 256, 0, (cudaStream_t)stream>>>(d_params, d_grads, d_exp_avg, d_exp_avg_sq,
                                                                         (long long)n, lr, beta1, beta2, eps,
                                                                         weight_decay, grad_scale, d_state2);
    PDR_CUDA_CHECK(cudaGetLastError());
    stage_mark((cudaStream_t)stream, ST_OPTIM);
    return PDEREAD_OK;
}

int pderead_sgd_step(float* d_params, const float* d_grads, float* d_momentum_buf, int64_t n, float lr, float momentum,
                     int nesterov, float weight_decay, float grad_scale, float* d_state2, void* stream) {
    if (!d_params || !d_grads || !d_state2 || n < 0 || (momentum != 0.f && !d_momentum_buf)) return PDEREAD_ERR_BAD_ARGUMENT;
    if (n == 0) return PDEREAD_OK;
    { const int rc_ = check_device(); if (rc_ != PDEREAD_OK) return rc_; }
    sgd_step_kernel<<<optim_blocks(n), 256, 0, (cudaStream_t)stream>>>(d_params, d_grads, d_momentum_buf, (long long)n,
                                                                        lr, momentum, nesterov, weight_decay,
                                                                        grad_scale, d_state2);
    PDR_CUDA_CHECK(cudaGetLastError());
    stage_mark((cudaStream_t)stream, ST_OPTIM);
    return PDEREAD_OK;
}

// ------------------------------------------------------------------------------------------
// sharded step: loss + point count in the tail of the gradient all-reduce
// ------------------------------------------------------------------------------------------
int pderead_loss_tail_pack(const double* d_sum_sq, int64_t M_local, float* d_tail4, void* stream) {
    if (!d_sum_sq || !d_tail4 || M_local < 0) return PDEREAD_ERR_BAD_ARGUMENT;
    { const int rc_ = check_device(); if (rc_ != PDEREAD_OK) return rc_; }
    loss_tail_pack_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_sum_sq, (long long)M_local, d_tail4);
    PDR_CUDA_CHECK(cudaGetLastError());
    return PDEREAD_OK;
}

int pderead_loss_tail_unpack(const float* d_tail4, int64_t M_assumed, double* d_out3, void* stream) {
    if (!d_tail4 || !d_out3 || M_assumed < 0) return PDEREAD_ERR_BAD_ARGUMENT;
    { const int rc_ = check_device(); if (rc_ != PDEREAD_OK) return rc_; }
    loss_tail_unpack_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_tail4, (double)M_assumed, d_out3);
    PDR_CUDA_CHECK(cudaGetLastError());
    return PDEREAD_OK;
}

// ------------------------------------------------------------------------------------------
// points, probe
// ------------------------------------------------------------------------------------------
int pderead_generate_points(const double* h_bounds, int dim, int64_t num_points, uint64_t seed, uint64_t offset,
                            float* d_points, void* stream) {
    if (!h_bounds || dim < 1 || dim > MAXD || num_points < 0 || (num_points > 0 && !d_points))
        return PDEREAD_ERR_BAD_ARGUMENT;
    if (num_points == 0) return PDEREAD_OK;
    { const int rc_ = check_device(); if (rc_ != PDEREAD_OK) return rc_; }
    BoundsDev bd;
    for (int d = 0; d < dim; ++d) {
        if (h_bounds[2 * d] > h_bounds[2 * d + 1]) return PDEREAD_ERR_BAD_ARGUMENT;   // reference Points.py:45-46
        bd.lo[d] = (float)h_bounds[2 * d];
        bd.span[d] = (float)(h_bounds[2 * d + 1] - h_bounds[2 * d]);
    }
    const long long blocks = (num_points + 255) / 256;
    generate_points_kernel<<<(int)(blocks < 2368 ? blocks : 2368), 256, 0, (cudaStream_t)stream>>>(
        bd, dim, num_points, seed, offset, d_points);
    PDR_CUDA_CHECK(cudaGetLastError());
    return PDEREAD_OK;
}

int pderead_fma_peak_probe(int variant, int iters, float* d_sink, double* h_flops, void* stream) {
    if (!d_sink || iters < 1 || variant < 0 || variant > 2) return PDEREAD_ERR_BAD_ARGUMENT;
    { const int rc_ = check_device(); if (rc_ != PDEREAD_OK) return rc_; }
    int dev = 0, sms = 0;
    PDR_CUDA_CHECK(cudaGetDevice(&dev));
    PDR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, threads = 256;
    fma_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(variant, iters, d_sink);
    PDR_CUDA_CHECK(cudaGetLastError());
    if (h_flops) *h_flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
    return PDEREAD_OK;
}

}  // extern "C"
