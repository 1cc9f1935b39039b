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
constexpr int GB = 64;   // edge of the Gram block one CTA accumulates
constexpr int GP = 32;   // points per tile
constexpr int GS = GP + 1;

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

// LIB[c*GP + p] for one tile of GP points; D[(d)*GP + p] holds D_x^d U; valid[p] masks the tail.
__device__ __forceinline__ void build_library_tile(const LibTable& t, const float* D, float* LIB, const float* mask) {
    for (int idx = threadIdx.x; idx < GP; idx += blockDim.x) LIB[idx] = mask[idx];
    __syncthreads();
    for (int k = 1; k <= t.K; ++k) {
        const int lo = t.level_off[k], hi = t.level_off[k + 1];
        for (int idx = threadIdx.x; idx < (hi - lo) * GP; idx += blockDim.x) {
            const int c = lo + idx / GP, p = idx % GP;
            LIB[c * GP + p] = LIB[t.parent[c] * GP + p] * D[t.last[c] * GP + p];
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
        build_library_tile(t, D, LIB, MASK);
        for (int idx = threadIdx.x; idx < GP * t.C; idx += blockDim.x) {
            const int p = idx / t.C, c = idx - p * t.C;
            if (p0 + p < M) lib[(p0 + p) * t.C + c] = LIB[c * GP + p];
        }
        __syncthreads();
    }
}

// Fused library + Gram.  Grid = (point slices, block pairs bi <= bj of the augmented Gram
// [A | b]^T [A | b]).  Each CTA keeps a 64x64 float64 block in registers (4x4 per thread) over
// all its tiles and adds it to the global result once at the end.
template <bool DENSE>
__global__ void __launch_bounds__(256) gram_kernel(LibTable t, const float* __restrict__ src, const float* __restrict__ bvec,
                                                   long long M, int nblk, double* __restrict__ G,
                                                   double* __restrict__ Atb, double* __restrict__ btb) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int C = t.C, CA = C + 1;
    const int nd = t.n + 1;
    double* Ai = reinterpret_cast<double*>(smraw);   // [GB][GS]
    double* Aj = Ai + GB * GS;                       // [GB][GS]
    float* LIB = reinterpret_cast<float*>(Aj + GB * GS);   // [C][GP]
    float* D = LIB + (size_t)C * GP;                 // [(n+1)][GP]
    float* MASK = D + nd * GP;                       // [GP]
    float* BV = MASK + GP;                           // [GP]

    // decode block pair
    int bi = 0, bj = 0;
    {
        int rem = blockIdx.y;
        for (bi = 0; bi < nblk; ++bi) {
            const int cnt = nblk - bi;
            if (rem < cnt) { bj = bi + rem; break; }
            rem -= cnt;
        }
    }
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;

    const long long tiles = (M + GP - 1) / GP;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long p0 = tile * GP;
        for (int idx = threadIdx.x; idx < GP; idx += blockDim.x) {
            const bool ok = (p0 + idx < M);
            MASK[idx] = ok ? 1.f : 0.f;
            BV[idx] = ok ? bvec[p0 + idx] : 0.f;
        }
        if (DENSE) {
            for (int idx = threadIdx.x; idx < GP * C; idx += blockDim.x) {
                const int p = idx / C, c = idx - p * C;
                LIB[c * GP + p] = (p0 + p < M) ? src[(p0 + p) * C + c] : 0.f;
            }
            __syncthreads();
        } else {
            for (int idx = threadIdx.x; idx < nd * GP; idx += blockDim.x) {
                const int p = idx / nd, d = idx - p * nd;
                D[d * GP + p] = (p0 + p < M) ? src[(p0 + p) * nd + d] : 0.f;
            }
            __syncthreads();
            build_library_tile(t, D, LIB, MASK);
        }
        // widen the two column blocks to float64
        for (int idx = threadIdx.x; idx < GB * GP; idx += blockDim.x) {
            const int r = idx / GP, p = idx - r * GP;
            const int ci = bi * GB + r, cj = bj * GB + r;
            Ai[r * GS + p] = (ci < C) ? (double)LIB[ci * GP + p] : (ci == C ? (double)BV[p] : 0.0);
            Aj[r * GS + p] = (cj < C) ? (double)LIB[cj * GP + p] : (cj == C ? (double)BV[p] : 0.0);
        }
        __syncthreads();
#pragma unroll 4
        for (int p = 0; p < GP; ++p) {
            double av[4], bv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) av[a] = Ai[(ty + 16 * a) * GS + p];
#pragma unroll
            for (int b = 0; b < 4; ++b) bv[b] = Aj[(tx + 16 * b) * GS + p];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int i = bi * GB + ty + 16 * a, j = bj * GB + tx + 16 * b;
            if (i >= CA || j >= CA) continue;
            const double v = acc[a][b];
            if (i < C && j < C) {
                atomicAdd(&G[(size_t)i * C + j], v);
                if (bi != bj) atomicAdd(&G[(size_t)j * C + i], v);
            } else if (i < C && j == C) {
                atomicAdd(&Atb[i], v);
            } else if (i == C && j == C) {
                atomicAdd(btb, v);
            }   // (i == C, j < C) duplicates (j, C) of the same diagonal block: skipped
        }
}

static size_t gram_smem_bytes(const LibTable& t) {
    return sizeof(double) * 2 * GB * GS + sizeof(float) * ((size_t)t.C * GP + (t.n + 1) * GP + 2 * GP);
}

int gram_launch(const LibTable& t, bool dense, const float* src, const float* b, long long M, double* G, double* Atb,
                double* btb, cudaStream_t s) {
    if (M <= 0) return PDEREAD_OK;
    const int CA = t.C + 1;
    const int nblk = (CA + GB - 1) / GB;
    const int npairs = nblk * (nblk + 1) / 2;
    int dev = 0, sms = 0;
    PDR_CUDA_CHECK(cudaGetDevice(&dev));
    PDR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long tiles = (M + GP - 1) / GP;
    long long slices = ((long long)sms * 2 + npairs - 1) / npairs;
    if (slices > tiles) slices = tiles;
    if (slices < 1) slices = 1;
    const size_t smem = gram_smem_bytes(t);
    dim3 grid((unsigned)slices, (unsigned)npairs);
    if (dense) {
        { const int rc_ = ensure_dynamic_smem((const void*)gram_kernel<true>, smem); if (rc_ != PDEREAD_OK) return rc_; }
        gram_kernel<true><<<grid, 256, smem, s>>>(t, src, b, M, nblk, G, Atb, btb);
    } else {
        { const int rc_ = ensure_dynamic_smem((const void*)gram_kernel<false>, smem); if (rc_ != PDEREAD_OK) return rc_; }
        gram_kernel<false><<<grid, 256, smem, s>>>(t, src, b, M, nblk, G, Atb, btb);
    }
    PDR_CUDA_CHECK(cudaGetLastError());
    stage_mark(s);
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
// RFE on the Gram system, one CTA.
//
// With unit-norm columns A_s = A diag(1/||A_i||) the normal equations of step j are
//   Gs[S,S] x = cs[S],   Gs = G / (norm norm^T),  cs = A^T b / norm,
// solved by Cholesky in float64 with the right-hand side carried as an extra row, so that
// w = L^-1 cs and ||b - A_s x||^2 = b^T b - w^T w.  (The reference solves the same least-squares
// problem with LAPACK dgelsd on the tall matrix, Extraction.py:464.)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rfe_kernel(const double* __restrict__ G, const double* __restrict__ Atb,
                                                  const double* __restrict__ btb, int C, float* __restrict__ X,
                                                  float* __restrict__ residual, int* __restrict__ order,
                                                  int* __restrict__ rank, float* __restrict__ change,
                                                  double* __restrict__ gws, int use_smem) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int LS = (C + 1) | 1;                  // odd row stride (doubles): conflict-free column walks
    double* Lm = use_smem ? reinterpret_cast<double*>(smraw) : gws;          // [(C+1)][LS]
    double* small = use_smem ? Lm + (size_t)(C + 1) * LS : reinterpret_cast<double*>(smraw);
    double* nrm = small;                         // [C]
    double* tmp = nrm + C;                       // [C+1]
    double* xs = tmp + C + 1;                    // [C]
    double* wv = xs + C;                         // [C]
    int* act = reinterpret_cast<int*>(wv + C);   // [C]
    int* alive = act + C;                        // [C]
    __shared__ int s_m, s_drop;
    __shared__ double s_piv;

    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < C; i += nt) {
        nrm[i] = sqrt(G[(size_t)i * C + i]);
        alive[i] = 1;
    }
    for (int i = tid; i < C * C; i += nt) X[i] = 0.f;
    __syncthreads();
    const double bb = btb[0];

    for (int step = 0; step < C; ++step) {
        if (tid == 0) {
            int m = 0;
            for (int i = 0; i < C; ++i)
                if (alive[i]) act[m++] = i;
            s_m = m;
        }
        __syncthreads();
        const int m = s_m;
        // scaled Gram of the active set (lower triangle) with the scaled right-hand side as row m
        for (int idx = tid; idx < (m + 1) * m; idx += nt) {
            const int a = idx / m, b = idx - a * m;
            if (a < m) {
                if (b <= a) {
                    const int ia = act[a], ib = act[b];
                    const double den = nrm[ia] * nrm[ib];
                    Lm[(size_t)a * LS + b] = den > 0.0 ? G[(size_t)ia * C + ib] / den : (a == b ? 1.0 : 0.0);
                }
            } else {
                const int ib = act[b];
                Lm[(size_t)m * LS + b] = nrm[ib] > 0.0 ? Atb[ib] / nrm[ib] : 0.0;
            }
        }
        __syncthreads();
        // left-looking Cholesky, rows k..m (row m = right-hand side)
        for (int k = 0; k < m; ++k) {
            for (int i = k + tid; i <= m; i += nt) {
                double s = Lm[(size_t)i * LS + k];
                const double* ri = Lm + (size_t)i * LS;
                const double* rk = Lm + (size_t)k * LS;
                for (int t2 = 0; t2 < k; ++t2) s = fma(-ri[t2], rk[t2], s);
                tmp[i] = s;
            }
            __syncthreads();
            if (tid == 0) {
                const double d2 = tmp[k];
                // the scaled Gram has a unit diagonal: a pivot this small means the column is
                // numerically dependent on the earlier ones -> freeze it at zero
                s_piv = (d2 > 1e-13) ? sqrt(d2) : 1e150;
            }
            __syncthreads();
            const double d = s_piv;
            for (int i = k + tid; i <= m; i += nt) Lm[(size_t)i * LS + k] = (i == k) ? d : tmp[i] / d;
            __syncthreads();
        }
        // w = row m;  back substitution L^T x = w
        for (int i = tid; i < m; i += nt) wv[i] = Lm[(size_t)m * LS + i];
        __syncthreads();
        double ww = 0.0;
        if (tid == 0)
            for (int i = 0; i < m; ++i) ww = fma(wv[i], wv[i], ww);
        for (int k = m - 1; k >= 0; --k) {
            if (tid == 0) xs[k] = wv[k] / Lm[(size_t)k * LS + k];
            __syncthreads();
            const double xk = xs[k];
            for (int t2 = tid; t2 < k; t2 += nt) wv[t2] = fma(-Lm[(size_t)k * LS + t2], xk, wv[t2]);
            __syncthreads();
        }
        // record the candidate, pick the feature to drop: smallest |x| in float32, first index wins
        // ties (reference Extraction.py:467-476 compares the float32 entries of X with strict '<')
        for (int a = tid; a < m; a += nt) {
            const int ia = act[a];
            const float xf = (float)xs[a];
            X[(size_t)ia * C + step] = xf / (float)nrm[ia];
        }
        if (tid == 0) {
            int best = 0;
            float bv = fabsf((float)xs[0]);
            for (int a = 1; a < m; ++a) {
                const float v = fabsf((float)xs[a]);
                if (v < bv) { bv = v; best = a; }
            }
            s_drop = act[best];
            alive[act[best]] = 0;
            order[step] = act[best];
            const double r2 = bb - ww;
            residual[step] = (float)sqrt(r2 > 0.0 ? r2 : 0.0);
        }
        __syncthreads();
    }
    if (tid == 0) {
        residual[C] = (float)sqrt(bb > 0.0 ? bb : 0.0);
        // Rank_Candidate_Solutions: relative jump, ascending selection sort carrying indices
        for (int i = 0; i < C; ++i) {
            change[i] = (residual[i + 1] - residual[i]) / residual[i];
            rank[i] = i;
        }
        for (int i = 0; i < C; ++i) {
            int k = i;
            for (int j = i; j < C; ++j)
                if (change[j] < change[k]) k = j;
            const float tc = change[i]; change[i] = change[k]; change[k] = tc;
            const int tr = rank[i]; rank[i] = rank[k]; rank[k] = tr;
        }
    }
}

static size_t rfe_small_bytes(int C) { return sizeof(double) * (4 * (size_t)C + 1) + sizeof(int) * 2 * (size_t)C + 64; }
static size_t rfe_matrix_bytes(int C) { return sizeof(double) * (size_t)(C + 1) * ((C + 1) | 1); }

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
    int dev = 0, maxsm = 0;
    PDR_CUDA_CHECK(cudaGetDevice(&dev));
    PDR_CUDA_CHECK(cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t small = rfe_small_bytes(C), mat = rfe_matrix_bytes(C);
    const int use_smem = (small + mat + 1024 <= (size_t)maxsm) ? 1 : 0;
    if (!use_smem && (!ws || ws_bytes < mat)) return PDEREAD_ERR_WORKSPACE_TOO_SMALL;
    const size_t smem = use_smem ? small + mat : small;
    { const int rc_ = ensure_dynamic_smem((const void*)rfe_kernel, smem); if (rc_ != PDEREAD_OK) return rc_; }
    rfe_kernel<<<1, 256, smem, (cudaStream_t)stream>>>(d_G, d_Atb, d_btb, C, d_X, d_residual, d_order, d_rank, d_change,
                                                        (double*)ws, use_smem);
    PDR_CUDA_CHECK(cudaGetLastError());
    stage_mark((cudaStream_t)stream);
    return PDEREAD_OK;
}

}  // extern "C"
