// Extraction-mode kernels: monomial library, fused library + Gram accumulation, and Recursive
// Feature Elimination + ranking on the Gram system (see include/pderead_b200.h).
//
// Reference behaviour being replaced (paths under /root/reference/Code):
//   Generate_Library             Extraction.py:178-378   (library columns, left-to-right fp32 products)
//   Recursive_Feature_Elimination Extraction.py:382-501  (unit-norm columns, lstsq, drop min |x|)
//   Rank_Candidate_Solutions     Extraction.py:505-572   (selection sort on relative residual jump)
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "internal.h"

namespace pdr {

constexpr int MAXC = PDEREAD_MAX_LIBRARY_COLUMNS;
constexpr int MAXK = 8;
constexpr int GP = 32;   // points per tile

// column c (degree k >= 1) = column parent[c] (its length k-1 prefix) times D[last[c]]: the same
// left-to-right product order as the reference's inner loop (Extraction.py:364-366).
struct LibTable {
    int C, K, n;
    int level_off[MAXK + 2];          // columns of degree k are [level_off[k], level_off[k+1])
    unsigned short parent[MAXC];
    unsigned char last[MAXC];
};

static int build_table(int n, int K, LibTable* t, int32_t* indices /* [C][K] or null */) {
    if (n < 0 || n > 7 || K < 0 || K > MAXK) return PDEREAD_ERR_BAD_ARGUMENT;
    memset(t, 0, sizeof(*t));
    t->n = n;
    t->K = K;
    std::vector<std::vector<int>> prev{{}}, all{{}};
    std::vector<int> prev_col{0};
    t->level_off[0] = 0;
    t->level_off[1] = 1;
    int C = 1;
    for (int k = 1; k <= K; ++k) {
        // lexicographic non-decreasing k-tuples = every (k-1)-tuple, in order, extended by v >= its last entry
        std::vector<std::vector<int>> cur;
        std::vector<int> cur_col;
        for (size_t a = 0; a < prev.size(); ++a) {
            const int lo = prev[a].empty() ? 0 : prev[a].back();
            for (int v = lo; v <= n; ++v) {
                if (C >= MAXC) return PDEREAD_ERR_BAD_ARGUMENT;
                std::vector<int> tup = prev[a];
                tup.push_back(v);
                t->parent[C] = (unsigned short)prev_col[a];
                t->last[C] = (unsigned char)v;
                cur.push_back(tup);
                cur_col.push_back(C);
                all.push_back(tup);
                ++C;
            }
        }
        t->level_off[k + 1] = C;
        prev.swap(cur);
        prev_col.swap(cur_col);
    }
    t->C = C;
    if (indices) {
        for (int c = 0; c < C; ++c)
            for (int j = 0; j < K; ++j) indices[c * K + j] = (j < (int)all[c].size()) ? all[c][j] : -1;
    }
    return PDEREAD_OK;
}

// LIB[c*TPN + p] for one tile of TPN points; D[(d)*TPN + p] holds D_x^d U; mask[p] masks the tail.
// If AD != nullptr every column is also written widened to float64 into the Gram kernel's layout
// AD[c / T][c % T][p] (group stride GST).
template <int TPN, int T>
__device__ __forceinline__ void build_library_tile(const LibTable& t, const float* D, float* LIB, const float* mask,
                                                   double* AD, int GST) {
    for (int idx = threadIdx.x; idx < TPN; idx += blockDim.x) {
        const float v = mask[idx];
        LIB[idx] = v;
        if (AD) AD[idx] = (double)v;               // column 0 = group 0, row 0
    }
    __syncthreads();
    for (int k = 1; k <= t.K; ++k) {
        const int lo = t.level_off[k], hi = t.level_off[k + 1];
        for (int idx = threadIdx.x; idx < (hi - lo) * TPN; idx += blockDim.x) {
            const int c = lo + idx / TPN, p = idx % TPN;
            const float v = LIB[t.parent[c] * TPN + p] * D[t.last[c] * TPN + p];
            LIB[c * TPN + p] = v;
            if (AD) {
                const int g = c / T, r = c - g * T;
                AD[(size_t)g * GST + r * TPN + p] = (double)v;
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) library_dense_kernel(LibTable t, const float* __restrict__ dxn, long long M,
                                                            float* __restrict__ lib) {
    extern __shared__ __align__(16) float sm[];
    float* D = sm;                       // [(n+1)][GP]
    float* MASK = D + (t.n + 1) * GP;    // [GP]
    float* LIB = MASK + GP;              // [C][GP]
    const int nd = t.n + 1;
    const long long tiles = (M + GP - 1) / GP;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long p0 = tile * GP;
        for (int idx = threadIdx.x; idx < nd * GP; idx += blockDim.x) {
            const int p = idx / nd, d = idx - p * nd;
            D[d * GP + p] = (p0 + p < M) ? dxn[(p0 + p) * nd + d] : 0.f;
        }
        for (int idx = threadIdx.x; idx < GP; idx += blockDim.x) MASK[idx] = 1.f;
        __syncthreads();
        build_library_tile<GP, 6>(t, D, LIB, MASK, nullptr, 0);
        for (int idx = threadIdx.x; idx < GP * t.C; idx += blockDim.x) {
            const int p = idx / t.C, c = idx - p * t.C;
            if (p0 + p < M) lib[(p0 + p) * t.C + c] = LIB[c * GP + p];
        }
        __syncthreads();
    }
}

// Fused library + Gram of the augmented matrix [A | b] (CA = C + 1 columns), upper triangle only.
//
// The columns are cut into groups of T = 6; a thread owns one T x T tile (gi <= gj) of [A|b]^T [A|b] in float64
// registers over all the point tiles of its CTA, so the symmetric half is never computed (C = 56: 1980 DFMA
// per point instead of the 4096 of a full 64 x 64 block).  A CTA covers `ntile_cta` consecutive tiles of the
// list (blockIdx.y selects the chunk; large libraries need several chunks for the register file) and SUB
// sub-slices of the points of a tile (warp-uniform), and adds its result to the global system once at the end.
// Point tiles are TPN = 128 points when the widened tile fits two CTAs per SM (the build / barrier phases between
// two compute phases are latency-bound: ncu put `barrier` at 2.6 stalls per issue with 32-point tiles), else 32.
// The derivatives of the NEXT tile are fetched into registers while the current one is being multiplied.
// Shared layout of the widened tile: AD[group][row in group][point], strides (T*TPN + 1, TPN, 1) doubles: the odd
// group stride spreads the groups a warp touches over the 16 8-byte banks.
constexpr int GT = 6;
constexpr int GRAM_PRE = 6;      // prefetch registers per thread: (n + 2) * TPN / threads <= 6

template <bool DENSE, int TPN>
__global__ void __launch_bounds__(256, 2) gram_kernel(LibTable t, const float* __restrict__ src,
                                                      const float* __restrict__ bvec, long long M, int ngroups,
                                                      int ntile_total, int ntile_cta, int sub, double* __restrict__ G,
                                                      double* __restrict__ Atb, double* __restrict__ btb) {
    extern __shared__ __align__(16) unsigned char smraw[];
    constexpr int T = GT;
    constexpr int GST = T * TPN + 1;
    const int C = t.C, CA = C + 1;
    const int nd = t.n + 1;
    double* AD = reinterpret_cast<double*>(smraw);                       // [ngroups][T][TPN] (+1 per group)
    float* LIB = reinterpret_cast<float*>(AD + (size_t)ngroups * GST);   // [C][TPN]
    float* D = LIB + (size_t)C * TPN;                                    // [(n+1)][TPN]
    float* MASK = D + nd * TPN;                                          // [TPN]
    float* BV = MASK + TPN;                                              // [TPN]

    // this thread's tile: index into the list of (gi <= gj) pairs, row-major over gi; the tile count is padded to a
    // warp so that a warp works on ONE sub-slice (measured: packing (sub-slice, tile) densely -- 220 of 224 lanes busy
    // at C = 56 instead of 220 of 256 -- is 3 % slower: lanes of different sub-slices read different points)
    const int tiles_pad = (ntile_cta + 31) & ~31;
    const int my_sub = threadIdx.x / tiles_pad;
    const int tl = threadIdx.x - my_sub * tiles_pad;
    const int tile_id = blockIdx.y * ntile_cta + tl;
    const bool has_tile = (tl < ntile_cta) && (tile_id < ntile_total) && (my_sub < sub);
    int gi = 0, gj = 0;
    if (has_tile) {
        int rem = tile_id;
        for (gi = 0; gi < ngroups; ++gi) {
            const int cnt = ngroups - gi;
            if (rem < cnt) { gj = gi + rem; break; }
            rem -= cnt;
        }
    }
    double acc[T][T];
#pragma unroll
    for (int a = 0; a < T; ++a)
#pragma unroll
        for (int b = 0; b < T; ++b) acc[a][b] = 0.0;

    // columns C+1 .. ngroups*T-1 of the last group are padding: zero once
    for (int idx = threadIdx.x; idx < (ngroups * T - CA) * TPN; idx += blockDim.x) {
        const int c = CA + idx / TPN, p = idx % TPN;
        AD[(size_t)(c / T) * GST + (c % T) * TPN + p] = 0.0;
    }

    const long long tiles = (M + TPN - 1) / TPN;
    const int nin = (nd + 1) * TPN;          // staged inputs of a tile: nd derivative rows + b
    float pre[GRAM_PRE];
    auto fetch = [&](long long tile) {
        const long long p0 = tile * TPN;
#pragma unroll
        for (int u = 0; u < GRAM_PRE; ++u) {
            const int idx = threadIdx.x + u * blockDim.x;
            float v = 0.f;
            if (idx < nin && tile < tiles) {
                const int p = idx / (nd + 1), d = idx - p * (nd + 1);
                if (p0 + p < M) v = (d < nd) ? __ldg(src + (p0 + p) * nd + d) : __ldg(bvec + p0 + p);
            }
            pre[u] = v;
        }
    };
    auto stage = [&](long long tile) {
        const long long p0 = tile * TPN;
#pragma unroll
        for (int u = 0; u < GRAM_PRE; ++u) {
            const int idx = threadIdx.x + u * blockDim.x;
            if (idx < nin) {
                const int p = idx / (nd + 1), d = idx - p * (nd + 1);
                if (d < nd) D[d * TPN + p] = pre[u];
                else BV[p] = pre[u];
            }
        }
        for (int idx = threadIdx.x; idx < TPN; idx += blockDim.x) MASK[idx] = (p0 + idx < M) ? 1.f : 0.f;
    };
    if (!DENSE) {
        fetch(blockIdx.x);
        stage(blockIdx.x);
    }
    __syncthreads();

    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long p0 = tile * TPN;
        if (DENSE) {
            for (int idx = threadIdx.x; idx < TPN; idx += blockDim.x) {
                const bool ok = (p0 + idx < M);
                BV[idx] = ok ? bvec[p0 + idx] : 0.f;
            }
            for (int idx = threadIdx.x; idx < TPN * C; idx += blockDim.x) {
                const int p = idx / C, c = idx - p * C;
                const float v = (p0 + p < M) ? src[(p0 + p) * C + c] : 0.f;
                const int g = c / T, r = c - g * T;
                AD[(size_t)g * GST + r * TPN + p] = (double)v;
            }
            __syncthreads();
        } else {
            build_library_tile<TPN, T>(t, D, LIB, MASK, AD, GST);
        }
        for (int idx = threadIdx.x; idx < TPN; idx += blockDim.x)
            AD[(size_t)(C / T) * GST + (C % T) * TPN + idx] = (double)BV[idx];
        __syncthreads();
        if (!DENSE) fetch(tile + gridDim.x);           // global loads in flight during the products
        if (has_tile) {
            const double* pa = AD + (size_t)gi * GST;
            const double* pb = AD + (size_t)gj * GST;
#pragma unroll 2
            for (int p = my_sub; p < TPN; p += sub) {
                double av[T], bv[T];
#pragma unroll
                for (int a = 0; a < T; ++a) av[a] = pa[a * TPN + p];
#pragma unroll
                for (int b = 0; b < T; ++b) bv[b] = pb[b * TPN + p];
#pragma unroll
                for (int a = 0; a < T; ++a)
#pragma unroll
                    for (int b = 0; b < T; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
            }
        }
        __syncthreads();
        if (!DENSE) {
            stage(tile + gridDim.x);
            __syncthreads();
        }
    }

    // combine the sub-slices in shared memory (reuses AD), then one float64 atomic per entry and CTA
    double* RED = AD;      // [T*T][tiles_pad] (host sizes the buffer for it)
    for (int s = 1; s < sub; ++s) {
        if (has_tile && my_sub == s) {
#pragma unroll
            for (int a = 0; a < T; ++a)
#pragma unroll
                for (int b = 0; b < T; ++b) RED[(size_t)(a * T + b) * tiles_pad + tl] = acc[a][b];
        }
        __syncthreads();
        if (has_tile && my_sub == 0) {
#pragma unroll
            for (int a = 0; a < T; ++a)
#pragma unroll
                for (int b = 0; b < T; ++b) acc[a][b] += RED[(size_t)(a * T + b) * tiles_pad + tl];
        }
        __syncthreads();
    }
    if (has_tile && my_sub == 0) {
#pragma unroll
        for (int a = 0; a < T; ++a)
#pragma unroll
            for (int b = 0; b < T; ++b) {
                const int i = gi * T + a, j = gj * T + b;
                if (i >= CA || j >= CA || j < i) continue;          // (diagonal tiles: upper triangle only)
                const double v = acc[a][b];
                if (j < C) {
                    atomicAdd(&G[(size_t)i * C + j], v);
                    if (i != j) atomicAdd(&G[(size_t)j * C + i], v);
                } else if (i < C) {
                    atomicAdd(&Atb[i], v);
                } else {
                    atomicAdd(btb, v);
                }
            }
    }
}

struct GramGeom {
    int tpn, ngroups, ntile, nchunk, ntile_cta, sub, threads;
    size_t smem;
};

static GramGeom gram_geometry(const LibTable& t, bool dense) {
    const int CA = t.C + 1;
    GramGeom g;
    memset(&g, 0, sizeof(g));
    g.ngroups = (CA + GT - 1) / GT;
    g.ntile = g.ngroups * (g.ngroups + 1) / 2;
    const int cap = 256;                                       // tiles per CTA (128 registers per thread, 2 CTAs per SM)
    g.nchunk = (g.ntile + cap - 1) / cap;
    g.ntile_cta = (g.ntile + g.nchunk - 1) / g.nchunk;
    const int pad = (g.ntile_cta + 31) & ~31;
    g.sub = cap / pad;
    if (g.sub < 1) g.sub = 1;
    if (g.sub > 8) g.sub = 8;
    g.threads = (pad * g.sub + 31) & ~31;
    if (g.threads < 64) g.threads = 64;                       // (the register prefetch needs (n + 2) * 32 / 6 threads)
    const int tpns[2] = {128, 32};
    for (int k = 0; k < 2; ++k) {
        g.tpn = tpns[k];
        const size_t ad = sizeof(double) * (size_t)g.ngroups * (GT * g.tpn + 1);
        const size_t red = g.sub > 1 ? sizeof(double) * (size_t)pad * GT * GT : 0;
        g.smem = (ad > red ? ad : red) + sizeof(float) * ((size_t)t.C * g.tpn + (t.n + 1) * g.tpn + 2 * g.tpn) + 16;
        // the register prefetch must cover a tile's inputs
        const bool pre_ok = dense || (size_t)(t.n + 2) * g.tpn <= (size_t)GRAM_PRE * g.threads;
        if (g.smem <= 110 * 1024 && pre_ok) break;
    }
    return g;
}

int gram_launch(const LibTable& t, bool dense, const float* src, const float* b, long long M, double* G, double* Atb,
                double* btb, cudaStream_t s) {
    if (M <= 0) return PDEREAD_OK;
    { const int rc_ = check_device(); if (rc_ != PDEREAD_OK) return rc_; }
    const GramGeom g = gram_geometry(t, dense);
    if (!dense && (size_t)(t.n + 2) * g.tpn > (size_t)GRAM_PRE * g.threads) return PDEREAD_ERR_UNSUPPORTED_ORDER;
    int dev = 0, sms = 0;
    PDR_CUDA_CHECK(cudaGetDevice(&dev));
    PDR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long tiles = (M + g.tpn - 1) / g.tpn;
    long long slices = ((long long)sms * 2 + g.nchunk - 1) / g.nchunk;
    if (slices > tiles) slices = tiles;
    if (slices < 1) slices = 1;
    dim3 grid((unsigned)slices, (unsigned)g.nchunk);
#define PDR_GRAM_LAUNCH(DENSE_, TPN_)                                                                                 \
    do {                                                                                                              \
        const int rc_ = ensure_dynamic_smem((const void*)gram_kernel<DENSE_, TPN_>, g.smem);                          \
        if (rc_ != PDEREAD_OK) return rc_;                                                                            \
        gram_kernel<DENSE_, TPN_><<<grid, g.threads, g.smem, s>>>(t, src, b, M, g.ngroups, g.ntile, g.ntile_cta,      \
                                                                  g.sub, G, Atb, btb);                                \
    } while (0)
    if (dense) {
        if (g.tpn == 128) PDR_GRAM_LAUNCH(true, 128); else PDR_GRAM_LAUNCH(true, 32);
    } else {
        if (g.tpn == 128) PDR_GRAM_LAUNCH(false, 128); else PDR_GRAM_LAUNCH(false, 32);
    }
#undef PDR_GRAM_LAUNCH
    PDR_CUDA_CHECK(cudaGetLastError());
    stage_mark(s, ST_GRAM);
    return PDEREAD_OK;
}

int library_gram_launch(const float* dxn, const float* b, long long M, int n, int K, double* G, double* Atb,
                        double* btb, cudaStream_t s) {
    LibTable t;
    int rc = build_table(n, K, &t, nullptr);
    if (rc) return rc;
    return gram_launch(t, false, dxn, b, M, G, Atb, btb, s);
}

// ---------------------------------------------------------------------------------------------
// RFE on the Gram system, one CTA, float64: the sweep operator.
//
// With unit-norm columns A_s = A diag(1/||A_i||) (reference Extraction.py:443-453) step j of the reference solves
//   min ||b - A_s[:, S] x||   (numpy.linalg.lstsq, Extraction.py:464)
// on the surviving set S.  On the augmented scaled Gram matrix  T = [[Gs, cs], [cs^T, b^T b]]  the sweep
//   SWP(k): d = T_kk;  T_ij -= T_ik T_kj / d (i, j != k);  T_ik = T_ki = T_ik / d;  T_kk = -1 / d
// applied to every k in S leaves  T_SS = -Gs_SS^-1,  T_S,rhs = x  and  T_rhs,rhs = ||b - A_s x||^2, and removing
// feature q from S is the reverse sweep, which on the kept entries is the same rank-1 update
//   T_ij -= T_iq T_qj / T_qq.
// So the kernel sweeps all C features in once (largest remaining diagonal first; a pivot below PIV_TOL means the
// column is numerically dependent on those already in: it is never swept, keeps coefficient 0 and leaves first)
// and then performs one rank-1 downdate per elimination: O(C^3) in total, every step fully parallel, two block
// barriers per step.  Ties in |x| go to the lowest index and the comparison is done on the float32 values, as in the
// reference (Extraction.py:467-476 scans the float32 X with a strict '<').
//
// Difference from the reference for rank-deficient libraries: dgelsd returns the minimum-norm solution (a
// duplicated column shares its coefficient with its twin and both survive until that share is the smallest
// entry); here the later twin is dropped at once with coefficient 0 and the eliminations that follow are those
// of the library without it.  A zero column gets coefficient 0 (the reference divides by its zero norm -> NaN).
// ---------------------------------------------------------------------------------------------
constexpr double RFE_PIV_TOL = 1e-13;
// RFE_MLP (template): independent matrix loads in flight per thread in the rank-1 updates: 1 when the matrix is in shared
// memory, 4 when it lives in global memory (C = 252: 8.7 -> 5.4 ms)

struct ArgBest {
    float v;
    int i;
};
// lexicographic (value, index) minimum over the warp
__device__ __forceinline__ ArgBest warp_argmin(ArgBest b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, b.v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, b.i, o);
        if (ov < b.v || (ov == b.v && oi < b.i)) { b.v = ov; b.i = oi; }
    }
    return b;
}
// every warp combines the per-warp partial results itself (no barrier, no serial section)
__device__ __forceinline__ ArgBest block_argmin_from_partials(const float* pv, const int* pi, int nwarps) {
    const int lane = threadIdx.x & 31;
    ArgBest b;
    b.v = lane < nwarps ? pv[lane] : INFINITY;
    b.i = lane < nwarps ? pi[lane] : 0x7fffffff;
    return warp_argmin(b);
}

template <int RFE_MLP>
__global__ void __launch_bounds__(1024) rfe_kernel(const double* __restrict__ G, const double* __restrict__ Atb,
                                                   const double* __restrict__ btb, int C, float* __restrict__ X,
                                                   float* __restrict__ residual, int* __restrict__ order,
                                                   int* __restrict__ rank, float* __restrict__ change,
                                                   double* __restrict__ gws, int use_smem) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int LD = C + 1;
    double* Tm = use_smem ? reinterpret_cast<double*>(smraw) : gws;          // [(C+1)][(C+1)]
    double* small = use_smem ? Tm + (size_t)LD * LD : reinterpret_cast<double*>(smraw);
    double* nrm = small;                         // [C]
    double* hv = nrm + C;                        // [C+1] pivot row of the current step
    double* dg = hv + C + 1;                     // [C]   running diagonal (forward phase)
    double* xs0 = dg + C;                        // [C]   coefficients of the active set, compact (double-buffered)
    double* xs1 = xs0 + C;
    int* act0 = reinterpret_cast<int*>(xs1 + C); // [C]   active features, ascending, compact (double-buffered)
    int* act1 = act0 + C;
    int* swept = act1 + C;                       // [C]
    float* pv = reinterpret_cast<float*>(swept + C);   // [32] per-warp partial results of the arg-min / arg-max
    int* pi = reinterpret_cast<int*>(pv + 32);         // [32]

    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const float inv_ld = 1.0f / (float)LD;

    for (int i = tid; i < C; i += nt) {
        const double g = G[(size_t)i * C + i];
        nrm[i] = g > 0.0 ? sqrt(g) : 0.0;
        swept[i] = 0;
    }
    for (int i = tid; i < C * C; i += nt) X[i] = 0.f;
    __syncthreads();
    for (int idx = tid; idx < LD * LD; idx += nt) {
        const int i = (int)(((float)idx + 0.5f) * inv_ld), j = idx - i * LD;
        double v;
        if (i < C && j < C) {
            const double den = nrm[i] * nrm[j];
            v = den > 0.0 ? G[(size_t)i * C + j] / den : 0.0;
        } else if (i < C || j < C) {
            const int c = i < C ? i : j;
            v = nrm[c] > 0.0 ? Atb[c] / nrm[c] : 0.0;
        } else {
            v = btb[0];
        }
        Tm[idx] = v;
    }
    const double bb = btb[0];
    // arg-max of the diagonal for the first pivot (negated: the helpers are arg-min)
    {
        ArgBest b;
        b.v = INFINITY; b.i = 0x7fffffff;
        for (int i = tid; i < C; i += nt) {
            const double d = nrm[i] > 0.0 ? 1.0 : 0.0;
            dg[i] = d;
            const float v = -(float)d;
            if (v < b.v || (v == b.v && i < b.i)) { b.v = v; b.i = i; }
        }
        b = warp_argmin(b);
        if (lane == 0) { pv[warp] = b.v; pi[warp] = b.i; }
    }
    __syncthreads();

    // ---- forward: sweep the features in, best-conditioned first ----
    int nswept = 0;
    for (int s = 0; s < C; ++s) {
        const ArgBest pb = block_argmin_from_partials(pv, pi, nwarps);
        const int k = pb.i;
        if (k < 0 || k >= C) break;
        const double d = dg[k];
        if (!(d >= RFE_PIV_TOL)) break;              // everything left is numerically dependent on the swept set
        for (int j = tid; j < LD; j += nt) hv[j] = Tm[(size_t)k * LD + j];
        __syncthreads();
        const double inv_d = 1.0 / d;
        // (batches of RFE_MLP independent loads, then the stores: when the matrix lives in global memory -- C > ~160 --
        //  a step is bound by the L2 round trip, and the compiler cannot hoist loads over stores to the same array)
        for (int base = tid; base < LD * LD; base += RFE_MLP * nt) {
            double cur[RFE_MLP];
#pragma unroll
            for (int u = 0; u < RFE_MLP; ++u) {
                const int idx = base + u * nt;
                cur[u] = idx < LD * LD ? Tm[idx] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < RFE_MLP; ++u) {
                const int idx = base + u * nt;
                if (idx < LD * LD) {
                    const int i = (int)(((float)idx + 0.5f) * inv_ld), j = idx - i * LD;
                    double v;
                    if (i == k) v = (j == k) ? -inv_d : hv[j] * inv_d;
                    else if (j == k) v = hv[i] * inv_d;
                    else v = fma(-hv[i] * inv_d, hv[j], cur[u]);
                    Tm[idx] = v;
                }
            }
        }
        ArgBest b;
        b.v = INFINITY; b.i = 0x7fffffff;
        for (int i = tid; i < C; i += nt) {
            if (i == k) swept[i] = 1;
            else if (!swept[i]) {
                const double nd2 = fma(-hv[i] * inv_d, hv[i], dg[i]);
                dg[i] = nd2;
                const float v = -(float)nd2;
                if (v < b.v || (v == b.v && i < b.i)) { b.v = v; b.i = i; }
            }
        }
        b = warp_argmin(b);
        if (lane == 0) { pv[warp] = b.v; pi[warp] = b.i; }   // (this step's partials were consumed before the barrier above)
        ++nswept;
        __syncthreads();
    }

    // ---- compact active list (ascending), coefficients, frozen list ----
    for (int i = tid; i < C; i += nt) {
        int pos = 0;
        for (int u = 0; u < i; ++u) pos += swept[u];
        if (swept[i]) {
            act0[pos] = i;
            xs0[pos] = Tm[(size_t)C * LD + i];
        } else {
            act1[i - pos] = i;                       // frozen features, ascending (act1 is free until the first swap)
        }
    }
    __syncthreads();
    const int nfrozen = C - nswept;
    double r2 = Tm[(size_t)C * LD + C];
    int m = nswept;
    int step = 0;
    // frozen features leave first, one per step, with the candidate solution unchanged
    for (; step < nfrozen; ++step) {
        for (int a = tid; a < m; a += nt) X[(size_t)act0[a] * C + step] = (float)xs0[a] / (float)nrm[act0[a]];
        if (tid == 0) {
            residual[step] = (float)sqrt(r2 > 0.0 ? r2 : 0.0);
            order[step] = act1[step];
        }
    }
    __syncthreads();
    // first arg-min over |x| (float32)
    {
        ArgBest b;
        b.v = INFINITY; b.i = 0x7fffffff;
        for (int a = tid; a < m; a += nt) {
            const float v = fabsf((float)xs0[a]);
            if (v < b.v || (v == b.v && a < b.i)) { b.v = v; b.i = a; }
        }
        b = warp_argmin(b);
        if (lane == 0) { pv[warp] = b.v; pi[warp] = b.i; }
    }
    __syncthreads();

    // ---- reverse: one rank-1 downdate per elimination ----
    double* xs = xs0; double* xn = xs1;
    int* act = act0; int* actn = act1;
    for (; step < C; ++step) {
        const ArgBest pb = block_argmin_from_partials(pv, pi, nwarps);
        const int aq = pb.i;                          // position of the feature to drop in the compact list
        const int q = act[aq];
        const double xq = xs[aq];
        const double d = Tm[(size_t)q * LD + q];      // = -(Gs_SS^-1)_qq < 0
        for (int a = tid; a < m; a += nt) {
            const int ia = act[a];
            hv[a] = Tm[(size_t)q * LD + ia];
            X[(size_t)ia * C + step] = (float)xs[a] / (float)nrm[ia];
        }
        if (tid == 0) {
            residual[step] = (float)sqrt(r2 > 0.0 ? r2 : 0.0);
            order[step] = q;
        }
        __syncthreads();
        const double inv_d = 1.0 / d;
        const float inv_m = 1.0f / (float)m;
        for (int base = tid; base < m * m; base += RFE_MLP * nt) {
            double cur[RFE_MLP];
            double* e[RFE_MLP];
            int ia[RFE_MLP], ib[RFE_MLP];
#pragma unroll
            for (int u = 0; u < RFE_MLP; ++u) {
                const int idx = base + u * nt;
                const bool ok = idx < m * m;
                const int a = ok ? (int)(((float)idx + 0.5f) * inv_m) : 0;
                ia[u] = a;
                ib[u] = ok ? idx - a * m : 0;
                e[u] = Tm + (size_t)act[ia[u]] * LD + act[ib[u]];
                cur[u] = ok ? *e[u] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < RFE_MLP; ++u)
                if (base + u * nt < m * m) *e[u] = fma(-hv[ia[u]] * inv_d, hv[ib[u]], cur[u]);
        }
        ArgBest b;
        b.v = INFINITY; b.i = 0x7fffffff;
        for (int a = tid; a < m; a += nt) {
            if (a == aq) continue;
            const int an = a - (a > aq ? 1 : 0);
            const double xv = fma(-hv[a] * inv_d, xq, xs[a]);
            xn[an] = xv;
            actn[an] = act[a];
            const float v = fabsf((float)xv);
            if (v < b.v || (v == b.v && an < b.i)) { b.v = v; b.i = an; }
        }
        b = warp_argmin(b);
        r2 = fma(-xq * inv_d, xq, r2);
        if (lane == 0) { pv[warp] = b.v; pi[warp] = b.i; }
        { double* tx = xs; xs = xn; xn = tx; }
        { int* ta = act; act = actn; actn = ta; }
        --m;
        __syncthreads();
    }

    // ---- Rank_Candidate_Solutions (Extraction.py:505-572): relative jump, ascending selection sort carrying the
    //      indices (swap with the first strict minimum of the tail), one warp, arg-min by shuffles ----
    if (tid == 0) residual[C] = (float)sqrt(bb > 0.0 ? bb : 0.0);
    __syncthreads();
    for (int i = tid; i < C; i += nt) {
        change[i] = (residual[i + 1] - residual[i]) / residual[i];
        rank[i] = i;
    }
    __syncthreads();
    if (warp == 0) {
        for (int i = 0; i < C; ++i) {
            ArgBest b;
            b.v = INFINITY; b.i = 0x7fffffff;
            bool any = false;
            for (int j = i + lane; j < C; j += 32) {
                const float v = change[j];
                if (!any || v < b.v) { b.v = v; b.i = j; any = true; }
            }
            if (!any) { b.v = INFINITY; b.i = 0x7fffffff; }
            b = warp_argmin(b);
            if (lane == 0 && b.i < C && b.i != i && change[b.i] < change[i]) {
                const float tc = change[i]; change[i] = change[b.i]; change[b.i] = tc;
                const int tr = rank[i]; rank[i] = rank[b.i]; rank[b.i] = tr;
            }
            __syncwarp();
        }
    }
}

static size_t rfe_small_bytes(int C) {
    return sizeof(double) * (5 * (size_t)C + 1) + sizeof(int) * 3 * (size_t)C + 64 * sizeof(float) + 64;
}
static size_t rfe_matrix_bytes(int C) { return sizeof(double) * (size_t)(C + 1) * (C + 1); }

}  // namespace pdr

using namespace pdr;

extern "C" {

int pderead_library_num_columns(int n, int K) {
    LibTable t;
    if (build_table(n, K, &t, nullptr) != PDEREAD_OK) return -1;
    return t.C;
}

int pderead_library_multi_indices(int n, int K, int32_t* h_indices) {
    LibTable t;
    if (!h_indices) return PDEREAD_ERR_BAD_ARGUMENT;
    return build_table(n, K, &t, h_indices);
}

int pderead_library_dense(const float* d_dxn, int64_t M, int n, int K, float* d_library, void* stream) {
    LibTable t;
    int rc = build_table(n, K, &t, nullptr);
    if (rc) return rc;
    if (M < 0 || (M > 0 && (!d_dxn || !d_library))) return PDEREAD_ERR_BAD_ARGUMENT;
    if (M == 0) return PDEREAD_OK;
    { const int rc_ = check_device(); if (rc_ != PDEREAD_OK) return rc_; }
    const size_t smem = sizeof(float) * ((size_t)(t.n + 1) * GP + GP + (size_t)t.C * GP);
    { const int rc_ = ensure_dynamic_smem((const void*)library_dense_kernel, smem); if (rc_ != PDEREAD_OK) return rc_; }
    const long long tiles = (M + GP - 1) / GP;
    library_dense_kernel<<<(int)(tiles < 1184 ? tiles : 1184), 256, smem, (cudaStream_t)stream>>>(t, d_dxn, M, d_library);
    PDR_CUDA_CHECK(cudaGetLastError());
    return PDEREAD_OK;
}

size_t pderead_library_gram_workspace_bytes(int, int) { return 0; }

int pderead_library_gram(const float* d_dxn, const float* d_b, int64_t M, int n, int K, double* d_G, double* d_Atb,
                         double* d_btb, void*, size_t, void* stream) {
    if (M < 0 || !d_G || !d_Atb || !d_btb || (M > 0 && (!d_dxn || !d_b))) return PDEREAD_ERR_BAD_ARGUMENT;
    return library_gram_launch(d_dxn, d_b, M, n, K, d_G, d_Atb, d_btb, (cudaStream_t)stream);
}

int pderead_dense_gram(const float* d_A, const float* d_b, int64_t M, int C, double* d_G, double* d_Atb, double* d_btb,
                       void*, size_t, void* stream) {
    if (M < 0 || C < 1 || C > MAXC || !d_G || !d_Atb || !d_btb || (M > 0 && (!d_A || !d_b))) return PDEREAD_ERR_BAD_ARGUMENT;
    LibTable t;
    memset(&t, 0, sizeof(t));
    t.C = C;
    return gram_launch(t, true, d_A, d_b, M, d_G, d_Atb, d_btb, (cudaStream_t)stream);
}

size_t pderead_rfe_workspace_bytes(int C) {
    if (C < 1 || C > MAXC) return 0;
    return rfe_matrix_bytes(C) + 256;
}

int pderead_rfe_gram(const double* d_G, const double* d_Atb, const double* d_btb, int C, float* d_X, float* d_residual,
                     int32_t* d_order, int32_t* d_rank, float* d_change, void* ws, size_t ws_bytes, void* stream) {
    if (C < 1 || C > MAXC || !d_G || !d_Atb || !d_btb || !d_X || !d_residual || !d_order || !d_rank || !d_change)
        return PDEREAD_ERR_BAD_ARGUMENT;
    { const int rc_ = check_device(); if (rc_ != PDEREAD_OK) return rc_; }
    int dev = 0, maxsm = 0;
    PDR_CUDA_CHECK(cudaGetDevice(&dev));
    PDR_CUDA_CHECK(cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t small = rfe_small_bytes(C), mat = rfe_matrix_bytes(C);
    const int use_smem = (small + mat + 1024 <= (size_t)maxsm) ? 1 : 0;
    if (!use_smem && (!ws || ws_bytes < mat)) return PDEREAD_ERR_WORKSPACE_TOO_SMALL;
    const size_t smem = use_smem ? small + mat : small;
    const void* kern = use_smem ? (const void*)rfe_kernel<1> : (const void*)rfe_kernel<4>;
    { const int rc_ = ensure_dynamic_smem(kern, smem); if (rc_ != PDEREAD_OK) return rc_; }
    int threads = (((C + 1) * (C + 1) / 8) + 31) & ~31;     // ~8 matrix entries per thread and step
    if (threads < 128) threads = 128;
    if (threads > 1024) threads = 1024;
    if (use_smem)
        rfe_kernel<1><<<1, threads, smem, (cudaStream_t)stream>>>(d_G, d_Atb, d_btb, C, d_X, d_residual, d_order, d_rank,
                                                                   d_change, (double*)ws, 1);
    else
        rfe_kernel<4><<<1, threads, smem, (cudaStream_t)stream>>>(d_G, d_Atb, d_btb, C, d_X, d_residual, d_order, d_rank,
                                                                   d_change, (double*)ws, 0);
    PDR_CUDA_CHECK(cudaGetLastError());
    stage_mark((cudaStream_t)stream, ST_RFE);
    return PDEREAD_OK;
}

}  // extern "C"
