// Register-tiled forward kernel of the plain MLP (jet length J = 1): the PDE network N(U, D_x U, ..), and U itself in
// Data_Loss (reference Network.py:174-202 called from PDE_Residual.py:211 and Loss_Functions.py:115-122).
//
// Why a second mapping for J = 1.  In mlp_jet_kernels.cuh a lane owns points and a warp owns 10 neurons; with one jet
// coefficient per point a k-step of the layer product is 3 uniform weight loads for 10 (or 20) FFMA2, and ncu shows the
// plain-MLP launches waiting on those loads (long_scoreboard 10 stalls per issue at width 100, FMA pipe 24 % busy).
// Here the layer's weights are staged in shared memory once per tile by cp.async (the next layer's copy overlaps the
// current layer's products when two buffers fit) and a thread owns an 8-neuron x PP-point register tile:
//   per k-step: 2 warp-uniform LDS.128 (8 weights) + 1 LDS.128 / LDS.64 (PP activations) -> 4 PP FFMA2.
// warp = group of 8 output neurons, lane = PP consecutive points of the tile (TP = 32 PP points: 128 for forward-only
// launches, 64 when the pre-activations are stashed for mlp_jet_bwd_kernel, whose tile is 64 points).
#pragma once
#include <cuda_runtime.h>

#include "internal.h"
#include "jet_math.cuh"
#include "mlp_jet_kernels.cuh"

namespace pdr {

constexpr int PG = 8;                 // neurons per warp
constexpr int PLAIN_MAX_WARPS = 16;   // widths up to 128
// Narrow networks (<= 7 warps, widths up to 56) get their own instantiation with a register cap that lets a third CTA
// share the SM (width 50: 7 warps, 70 KB of shared memory per CTA; at ptxas' free choice of 99 / 128 registers only two
// CTAs were resident and `barrier` was the top stall of the reverse kernel).
constexpr int PLAIN_NARROW_WARPS = 7;
#ifndef PDR_PLAIN_NARROW_CTAS
#define PDR_PLAIN_NARROW_CTAS 3
#endif
template <bool NARROW>
struct PlainBounds {
    static constexpr int threads = (NARROW ? PLAIN_NARROW_WARPS : PLAIN_MAX_WARPS) * 32;
    static constexpr int ctas = NARROW ? PDR_PLAIN_NARROW_CTAS : 1;
};

template <int PP>
__device__ __forceinline__ void vec_load(const float* p, float* v) {
    if constexpr (PP == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else if constexpr (PP == 2) {
        const float2 t = *reinterpret_cast<const float2*>(p);
        v[0] = t.x; v[1] = t.y;
    } else {
#pragma unroll
        for (int q = 0; q < PP; ++q) v[q] = p[q];
    }
}
template <int PP>
__device__ __forceinline__ void vec_store(float* p, const float* v) {
    if constexpr (PP == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else if constexpr (PP == 2) {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    } else {
#pragma unroll
        for (int q = 0; q < PP; ++q) p[q] = v[q];
    }
}

// a.wt_packed: [L-1][W][WP] with wt[l][k][i] = W_{l+1}[i][k] (zero for i >= W), WP = 8 * ceil(W / 8)
// a.aux: number of weight buffers in shared memory (1 or 2)
template <int ACT, bool STASH, int PP, bool NARROW>
__global__ void __launch_bounds__(PlainBounds<NARROW>::threads, PlainBounds<NARROW>::ctas) mlp_plain_fwd_kernel(const FwdArgs a) {
    constexpr int TP = 32 * PP;
    constexpr int HS = TP + 4;

    extern __shared__ __align__(16) float smem[];
    const NetDev& net = a.net;
    const int W = net.W, L = net.L, din = net.din;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int WP = nwarps * PG;
    const int nbuf = a.aux;

    float* Hbuf0 = smem;                          // [W][HS]
    float* Hbuf1 = Hbuf0 + (size_t)W * HS;
    float* IN = Hbuf1 + (size_t)W * HS;           // [din][TP]
    float* RED = IN + din * TP;                   // [nwarps][TP]
    float* WS = RED + nwarps * TP;                // [nbuf][W][WP]

    const int i0 = warp * PG, pt0 = lane * PP;
    double lsum = 0.0;
    const float* wout = net.w[L];

    auto stage_weights = [&](int l, int buf) {    // weights of hidden layer l (1 .. L-1) -> WS[buf]
        const float4* src = reinterpret_cast<const float4*>(a.wt_packed + (size_t)(l - 1) * W * WP);
        float4* dst = reinterpret_cast<float4*>(WS + (size_t)buf * W * WP);
        for (int i = threadIdx.x; i < W * WP / 4; i += blockDim.x) cp_async16(dst + i, src + i);
        cp_async_commit();
    };

    for (long long tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const long long p0 = tile * TP;
        if (L > 1) stage_weights(1, 0);            // (both buffers are free: the previous tile ended with a barrier)
        for (int idx = threadIdx.x; idx < din * TP; idx += blockDim.x) {
            const int pt = idx / din, d = idx - pt * din;
            const long long p = p0 + pt;
            IN[d * TP + pt] = (p < a.M) ? __ldg(a.in + p * din + d) : 0.f;
        }
        __syncthreads();

        float part[PP];
#pragma unroll
        for (int q = 0; q < PP; ++q) part[q] = 0.f;
        float* Hin = Hbuf0;
        float* Hout = Hbuf1;
        float ab[7];

        // ---- hidden layer 0: K = din is tiny, no staging ----
        load_ab<ACT>(net, 0, ab);
#pragma unroll 1
        for (int n = 0; n < PG; ++n) {
            const int i = i0 + n;
            if (i >= W) break;
            const float* wrow = net.w[0] + (size_t)i * din;
            float s[PP], y[PP];
            const float bv = __ldg(net.b[0] + i);
#pragma unroll
            for (int q = 0; q < PP; ++q) s[q] = bv;
            for (int d = 0; d < din; ++d) {
                const float wd = __ldg(wrow + d);
                float xin[PP];
                vec_load<PP>(IN + d * TP + pt0, xin);
#pragma unroll
                for (int q = 0; q < PP; ++q) s[q] = fmaf(wd, xin[q], s[q]);
            }
#pragma unroll
            for (int q = 0; q < PP; ++q) act_fwd<ACT, 0, 0, float>(&s[q], ab, &y[q]);
            if (L == 1) {
                const float wo = __ldg(wout + i);
#pragma unroll
                for (int q = 0; q < PP; ++q) part[q] = fmaf(wo, y[q], part[q]);
            } else {
                vec_store<PP>(Hin + (size_t)i * HS + pt0, y);
            }
        }
        cp_async_wait_all();
        __syncthreads();

        // ---- hidden layers 1 .. L-1 ----
        for (int l = 1; l < L; ++l) {
            const int buf = (nbuf == 2) ? ((l - 1) & 1) : 0;
            if (nbuf == 2 && l + 1 < L) stage_weights(l + 1, l & 1);
            const bool last = (l == L - 1);
            load_ab<ACT>(net, l, ab);
            float2 acc[PG / 2][PP];
            {
                const float* bias = net.b[l];
#pragma unroll
                for (int n = 0; n < PG; n += 2) {
                    const float b0 = (i0 + n < W) ? __ldg(bias + i0 + n) : 0.f;
                    const float b1 = (i0 + n + 1 < W) ? __ldg(bias + i0 + n + 1) : 0.f;
#pragma unroll
                    for (int q = 0; q < PP; ++q) acc[n / 2][q] = make_float2(b0, b1);
                }
            }
            {
                const float* wk = WS + (size_t)buf * W * WP + i0;
                const float* hk = Hin + pt0;
#pragma unroll 4
                for (int k = 0; k < W; ++k) {
                    const float4 w0 = *reinterpret_cast<const float4*>(wk);
                    const float4 w1 = *reinterpret_cast<const float4*>(wk + 4);
                    float hv[PP];
                    vec_load<PP>(hk, hv);
#pragma unroll
                    for (int q = 0; q < PP; ++q) {
                        const float2 a2 = make_float2(hv[q], hv[q]);
                        acc[0][q] = __ffma2_rn(a2, make_float2(w0.x, w0.y), acc[0][q]);
                        acc[1][q] = __ffma2_rn(a2, make_float2(w0.z, w0.w), acc[1][q]);
                        acc[2][q] = __ffma2_rn(a2, make_float2(w1.x, w1.y), acc[2][q]);
                        acc[3][q] = __ffma2_rn(a2, make_float2(w1.z, w1.w), acc[3][q]);
                    }
                    wk += WP;
                    hk += HS;
                }
            }
            if (nbuf == 1) {
                __syncthreads();                       // every warp is done with WS[0]
                if (l + 1 < L) stage_weights(l + 1, 0);
            }
            // epilogue: activation, stash, next layer's input (or the output dot product)
#pragma unroll
            for (int n = 0; n < PG; ++n) {
                const int i = i0 + n;
                if (i < W) {
                    float z[PP], y[PP];
#pragma unroll
                    for (int q = 0; q < PP; ++q) {
                        z[q] = (n & 1) ? acc[n / 2][q].y : acc[n / 2][q].x;
                        act_fwd<ACT, 0, 0, float>(&z[q], ab, &y[q]);
                    }
                    if (STASH) {
                        // layout of mlp_jet_bwd_kernel at J = 1: [tile][layer][neuron][point]; tanh rows carry y
                        float* sp = a.stash + (((size_t)tile * (L - 1) + (l - 1)) * W + i) * TP + pt0;
                        vec_store<PP>(sp, ACT == ACT_TANH ? y : z);
                    }
                    if (last) {
                        const float wo = __ldg(wout + i);
#pragma unroll
                        for (int q = 0; q < PP; ++q) part[q] = fmaf(wo, y[q], part[q]);
                    } else {
                        vec_store<PP>(Hout + (size_t)i * HS + pt0, y);
                    }
                }
            }
            cp_async_wait_all();
            __syncthreads();
            float* t = Hin; Hin = Hout; Hout = t;
        }

        // ---- output layer: sum the per-warp partial dot products ----
        vec_store<PP>(RED + warp * TP + pt0, part);
        __syncthreads();
        const float bout = __ldg(net.b[L]);
        for (int pt = threadIdx.x; pt < TP; pt += blockDim.x) {
            const long long p = p0 + pt;
            float u = bout;
            for (int w = 0; w < nwarps; ++w) u += RED[w * TP + pt];
            if (p >= a.M) continue;
            if (a.dtm_in != nullptr) {
                const float R = __ldg(a.dtm_in + p) - u;
                if (a.resid_out) a.resid_out[p] = R;
                if (a.seed_out) a.seed_out[p] = (a.grad_resid != nullptr) ? -__ldg(a.grad_resid + p) : -a.seed_scale * R;
                lsum += (double)R * (double)R;
            } else {
                a.out0[p] = u;
            }
        }
        __syncthreads();
    }

    if (a.sumsq != nullptr) {
        __shared__ double red_d[PLAIN_MAX_WARPS];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
        if (lane == 0) red_d[warp] = lsum;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < nwarps; ++w) s += red_d[w];
            atomicAdd(a.sumsq, s);
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
inline int plain_wp(int W) { return (W + PG - 1) / PG * PG; }
inline bool plain_supported(int W) { return W >= 1 && (W + PG - 1) / PG <= PLAIN_MAX_WARPS; }
inline size_t plain_smem_bytes(int W, int L, int din, int PP, int nbuf) {
    const int TP = 32 * PP, HS = TP + 4, nw = (W + PG - 1) / PG;
    return sizeof(float) * ((size_t)2 * W * HS + (size_t)din * TP + (size_t)nw * TP +
                            (L > 1 ? (size_t)nbuf * W * plain_wp(W) : 0));
}

template <int ACT, bool STASH, int PP>
int launch_plain_tile(FwdArgs a, cudaStream_t stream) {
    const int W = a.net.W, L = a.net.L, din = a.net.din;
    const int nw = (W + PG - 1) / PG;
    // two weight buffers unless that costs a resident CTA
    const size_t s1 = plain_smem_bytes(W, L, din, PP, 1), s2 = plain_smem_bytes(W, L, din, PP, 2);
    const size_t sm_bytes = 227 * 1024;
    const int c1 = (int)(sm_bytes / (s1 + 1024)), c2 = (int)(sm_bytes / (s2 + 1024));
    a.aux = (c2 >= 1 && (c2 == c1 || c2 >= 3)) ? 2 : 1;
    const size_t smem = a.aux == 2 ? s2 : s1;
    a.num_tiles = (a.M + 32 * PP - 1) / (32 * PP);
    const bool narrow = nw <= PLAIN_NARROW_WARPS;
    const void* k = narrow ? (const void*)mlp_plain_fwd_kernel<ACT, STASH, PP, true> : (const void*)mlp_plain_fwd_kernel<ACT, STASH, PP, false>;
    int grid = 0;
    int rc = grid_for(k, nw * 32, smem, a.num_tiles, &grid);
    if (rc != PDEREAD_OK) return rc;
    if (narrow) mlp_plain_fwd_kernel<ACT, STASH, PP, true><<<grid, nw * 32, smem, stream>>>(a);
    else        mlp_plain_fwd_kernel<ACT, STASH, PP, false><<<grid, nw * 32, smem, stream>>>(a);
    PDR_CUDA_CHECK(cudaGetLastError());
    return PDEREAD_OK;
}

template <int ACT>
int launch_plain_fwd(const FwdArgs& a, cudaStream_t stream) {
    // the stash tile is the tile of the reverse kernel that will read it: 64 points (mlp_jet_bwd_kernel at J = 1) or
    // 128 points (mlp_plain_bwd_kernel), a.aux2 says which
    if (a.stash) return a.aux2 == 4 ? launch_plain_tile<ACT, true, 4>(a, stream) : launch_plain_tile<ACT, true, 2>(a, stream);
    return launch_plain_tile<ACT, false, 4>(a, stream);
}

// =====================================================================================================================
// reverse pass of the plain MLP (J = 1), same thread mapping: warp = 8 neurons, lane = 4 points, tile = 128 points.
//   per hidden layer lv (top first), with Zb = zbar of the layer above and Yb = y of this layer in shared memory:
//   (b) gW_l += Zb Yb^T (+ bias column)            -- wgrad_tiles of mlp_jet_kernels.cuh (2-wavefront LDS.128 tiles)
//   (c) ybar = W_l^T Zb  as an 8 x 4 register tile with W_l staged in shared memory by cp.async (row i of the native
//       [out][in] matrix holds the 8 weights a warp needs: 2 warp-uniform LDS.128 + 1 LDS.128 of Zb per 16 FFMA2),
//       activation VJP in the epilogue, zbar overwrites y in Yb
//   (d) Zb <- y of the layer below from the stash, for the next weight gradient.
// a.wn_packed: [L-1][W][WP] with wn[l][i][k] = W_{l+1}[i][k] (zero for k >= W); a.aux = weight buffers (1 or 2)
// =====================================================================================================================
template <int ACT, bool NARROW>
__global__ void __launch_bounds__(PlainBounds<NARROW>::threads, PlainBounds<NARROW>::ctas) mlp_plain_bwd_kernel(const BwdArgs a) {
    constexpr int PP = 4, TP = 32 * PP, HS = TP + 4;

    extern __shared__ __align__(16) float smem[];
    const NetDev& net = a.net;
    const int W = net.W, L = net.L, din = net.din;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int WP = nwarps * PG;
    const int nbuf = a.aux;

    float* Zb = smem;
    float* Yb = Zb + (size_t)W * HS;
    float* IN = Yb + (size_t)W * HS;           // [din][TP]
    float* SD = IN + din * TP;                 // [TP] seeds d(loss)/d(out)
    float* GIN = SD + TP;                      // [din][TP]
    float* SACC = GIN + din * TP;              // small-parameter gradient accumulators (persist over tiles)
    const int sbL = L * W, sw0 = sbL + 1, swout = sw0 + W * din, sab = swout + W, NS = sab + 7 * L;
    float* ONES = SACC + ((NS + 3) & ~3);      // [HS]: 1 over the points (bias column of the weight-gradient tiles)
    float* WS = ONES + HS;                     // [nbuf][W][WP]

    const int i0 = warp * PG, pt0 = lane * PP;
    float* gp = a.gpart + (size_t)blockIdx.x * a.lay.total;
    if (!a.accumulate)
        for (int idx = threadIdx.x; idx < a.lay.total; idx += blockDim.x) gp[idx] = 0.f;
    for (int idx = threadIdx.x; idx < NS; idx += blockDim.x) SACC[idx] = 0.f;
    for (int idx = threadIdx.x; idx < HS; idx += blockDim.x) ONES[idx] = idx < TP ? 1.f : 0.f;
    __syncthreads();

    const float* wout = net.w[L];
    auto stage_weights = [&](int l, int buf) {    // matrix between hidden layers l-1 and l (1 .. L-1) -> WS[buf]
        const float4* src = reinterpret_cast<const float4*>(a.wn_packed + (size_t)(l - 1) * W * WP);
        float4* dst = reinterpret_cast<float4*>(WS + (size_t)buf * W * WP);
        for (int i = threadIdx.x; i < W * WP / 4; i += blockDim.x) cp_async16(dst + i, src + i);
        cp_async_commit();
    };
    // pre-activation row ("row": tanh rows hold y) of neuron i of hidden layer lv for this thread's PP points
    auto load_row = [&](int lv, long long tile, int i, float* z) {
        if (lv == 0) {
#pragma unroll
            for (int q = 0; q < PP; ++q) input_layer_row<ACT, 0, 0>(net, IN, TP, i, pt0 + q, &z[q]);
        } else {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(
                a.stash + (((size_t)tile * (L - 1) + (lv - 1)) * W + i) * TP + pt0));
            z[0] = v.x; z[1] = v.y; z[2] = v.z; z[3] = v.w;
        }
    };

    for (long long tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const long long p0 = tile * TP;
        if (L > 1) stage_weights(L - 1, 0);
        for (int idx = threadIdx.x; idx < din * TP; idx += blockDim.x) {
            const int pt = idx / din, d = idx - pt * din;
            const long long p = p0 + pt;
            IN[d * TP + pt] = (p < a.M) ? __ldg(a.in + p * din + d) : 0.f;
            GIN[d * TP + pt] = 0.f;
        }
        for (int pt = threadIdx.x; pt < TP; pt += blockDim.x) {
            const long long p = p0 + pt;
            SD[pt] = (p < a.M && a.g0) ? a.g0_scale * __ldg(a.g0 + p) : 0.f;
        }
        __syncthreads();

        float sd[PP];
        vec_load<PP>(SD + pt0, sd);
        if (warp == 0) {                       // output bias: sum of the seeds
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < PP; ++q) s += sd[q];
            s = warp_sum(s);
            if (lane == 0) SACC[sbL] += s;
        }

        // ---- top hidden layer lv = L-1: ybar = w_out * seed.  Zb <- zbar^{L-1}, Yb <- y^{L-2}
        {
            const int lv = L - 1;
            float ab[7], ab_lo[7], abbar[7];
            load_ab<ACT>(net, lv, ab);
            load_ab<ACT>(net, lv > 0 ? lv - 1 : 0, ab_lo);
#pragma unroll
            for (int t = 0; t < 7; ++t) abbar[t] = 0.f;
#pragma unroll 1
            for (int n = 0; n < PG; ++n) {
                const int i = i0 + n;
                if (i >= W) break;
                float z[PP], y[PP], zbar[PP];
                load_row(lv, tile, i, z);
                if (lv > 0) {
                    float zl[PP], yl[PP];
                    load_row(lv - 1, tile, i, zl);
#pragma unroll
                    for (int q = 0; q < PP; ++q) act_fwd_row<ACT, 0, 0>(&zl[q], ab_lo, &yl[q]);
                    vec_store<PP>(Yb + (size_t)i * HS + pt0, yl);
                }
                const float wo = __ldg(wout + i);
                float gwo = 0.f;
#pragma unroll
                for (int q = 0; q < PP; ++q) {
                    act_fwd_row<ACT, 0, 0>(&z[q], ab, &y[q]);
                    const float ybar = wo * sd[q];
                    gwo = fmaf(sd[q], y[q], gwo);
                    act_vjp_row<ACT, 0, 0>(&z[q], &y[q], ab, &ybar, &zbar[q], abbar);
                }
                vec_store<PP>(Zb + (size_t)i * HS + pt0, zbar);
                gwo = warp_sum(gwo);
                if (lane == 0) SACC[swout + i] += gwo;
            }
            if (ACT == ACT_RATIONAL) {
#pragma unroll
                for (int t = 0; t < 7; ++t) {
                    const float s = warp_sum(abbar[t]);
                    if (lane == 0) atomicAdd(&SACC[sab + 7 * lv + t], s);
                }
            }
            cp_async_wait_all();
            __syncthreads();
        }

        int buf = 0;
        for (int lv = L - 2; lv >= 0; --lv) {
            const int l = lv + 1;              // matrix between hidden layers lv and lv+1, in WS[buf]
            if (nbuf == 2 && l - 1 >= 1) stage_weights(l - 1, buf ^ 1);
            // (b) weight (and bias) gradient of matrix l
            {
                float* gw = gp + a.lay.off_w[l];
                if (a.wg_mode == 2)      wgrad_tiles<5, 4>(Zb, Yb, ONES, W, HS, TP, gw, SACC + l * W);
                else if (a.wg_mode == 1) wgrad_tiles<5, 3>(Zb, Yb, ONES, W, HS, TP, gw, SACC + l * W);
                else                     wgrad_tiles<5, 2>(Zb, Yb, ONES, W, HS, TP, gw, SACC + l * W);
            }
            __syncthreads();
            // (c) ybar^{lv}[k][p] = sum_i W_l[i][k] Zb[i][p] for this warp's 8 neurons k and this lane's 4 points
            float2 acc[PG / 2][PP];
#pragma unroll
            for (int m = 0; m < PG / 2; ++m)
#pragma unroll
                for (int q = 0; q < PP; ++q) acc[m][q] = make_float2(0.f, 0.f);
            {
                const float* wk = WS + (size_t)buf * W * WP + i0;
                const float* zk = Zb + pt0;
#pragma unroll 4
                for (int i = 0; i < W; ++i) {
                    const float4 w0 = *reinterpret_cast<const float4*>(wk);
                    const float4 w1 = *reinterpret_cast<const float4*>(wk + 4);
                    float zv[PP];
                    vec_load<PP>(zk, zv);
#pragma unroll
                    for (int q = 0; q < PP; ++q) {
                        const float2 a2 = make_float2(zv[q], zv[q]);
                        acc[0][q] = __ffma2_rn(a2, make_float2(w0.x, w0.y), acc[0][q]);
                        acc[1][q] = __ffma2_rn(a2, make_float2(w0.z, w0.w), acc[1][q]);
                        acc[2][q] = __ffma2_rn(a2, make_float2(w1.x, w1.y), acc[2][q]);
                        acc[3][q] = __ffma2_rn(a2, make_float2(w1.z, w1.w), acc[3][q]);
                    }
                    wk += WP;
                    zk += HS;
                }
            }
            if (nbuf == 1) {
                __syncthreads();               // every warp is done with WS[0]
                if (l - 1 >= 1) stage_weights(l - 1, 0);
            }
            // epilogue: activation VJP of layer lv; zbar^{lv} replaces y^{lv} in Yb (elements private to this thread)
            {
                float ab[7], abbar[7];
                load_ab<ACT>(net, lv, ab);
#pragma unroll
                for (int t = 0; t < 7; ++t) abbar[t] = 0.f;
#pragma unroll
                for (int n = 0; n < PG; ++n) {
                    const int k = i0 + n;
                    if (k < W) {
                        float z[PP], y[PP], zbar[PP];
                        load_row(lv, tile, k, z);
                        if (VJP_FROM_Y<ACT>::v) vec_load<PP>(Yb + (size_t)k * HS + pt0, y);
#pragma unroll
                        for (int q = 0; q < PP; ++q) {
                            const float ybar = (n & 1) ? acc[n / 2][q].y : acc[n / 2][q].x;
                            if (!VJP_FROM_Y<ACT>::v) y[q] = 0.f;
                            act_vjp_row<ACT, 0, 0>(&z[q], &y[q], ab, &ybar, &zbar[q], abbar);
                        }
                        vec_store<PP>(Yb + (size_t)k * HS + pt0, zbar);
                    }
                }
                if (ACT == ACT_RATIONAL) {
#pragma unroll
                    for (int t = 0; t < 7; ++t) {
                        const float s = warp_sum(abbar[t]);
                        if (lane == 0) atomicAdd(&SACC[sab + 7 * lv + t], s);
                    }
                }
            }
            cp_async_wait_all();
            __syncthreads();
            // (d) Zb (dead now) <- y^{lv-1} for the next weight gradient
            if (lv > 0) {
                float ab_lo[7];
                load_ab<ACT>(net, lv - 1, ab_lo);
#pragma unroll 2
                for (int n = 0; n < PG; ++n) {
                    const int i = i0 + n;
                    if (i < W) {
                        float zl[PP], yl[PP];
                        load_row(lv - 1, tile, i, zl);
#pragma unroll
                        for (int q = 0; q < PP; ++q) act_fwd_row<ACT, 0, 0>(&zl[q], ab_lo, &yl[q]);
                        vec_store<PP>(Zb + (size_t)i * HS + pt0, yl);
                    }
                }
                __syncthreads();
            }
            float* t = Zb; Zb = Yb; Yb = t;
            if (nbuf == 2) buf ^= 1;
        }

        // ---- input layer: Zb = zbar^0
        {
            float gi[PP][MAXD];
#pragma unroll
            for (int q = 0; q < PP; ++q)
#pragma unroll
                for (int d = 0; d < MAXD; ++d) gi[q][d] = 0.f;
#pragma unroll 1
            for (int n = 0; n < PG; ++n) {
                const int i = i0 + n;
                if (i >= W) break;
                const float* wrow = net.w[0] + (size_t)i * din;
                float zb[PP], gwin[MAXD], gb0 = 0.f;
                vec_load<PP>(Zb + (size_t)i * HS + pt0, zb);
#pragma unroll
                for (int d = 0; d < MAXD; ++d) gwin[d] = 0.f;
#pragma unroll
                for (int q = 0; q < PP; ++q) {
                    gb0 += zb[q];
#pragma unroll
                    for (int d = 0; d < MAXD; ++d)
                        if (d < din) {
                            gwin[d] = fmaf(zb[q], IN[d * TP + pt0 + q], gwin[d]);
                            gi[q][d] = fmaf(__ldg(wrow + d), zb[q], gi[q][d]);
                        }
                }
                gb0 = warp_sum(gb0);
#pragma unroll
                for (int d = 0; d < MAXD; ++d)
                    if (d < din) gwin[d] = warp_sum(gwin[d]);
                if (lane == 0) {
                    SACC[i] += gb0;
#pragma unroll
                    for (int d = 0; d < MAXD; ++d)
                        if (d < din) SACC[sw0 + i * din + d] += gwin[d];
                }
            }
            if (a.gin != nullptr) {
#pragma unroll
                for (int q = 0; q < PP; ++q)
#pragma unroll
                    for (int d = 0; d < MAXD; ++d)
                        if (d < din) atomicAdd(&GIN[d * TP + pt0 + q], gi[q][d]);
            }
        }
        __syncthreads();
        if (a.gin != nullptr) {
            for (int idx = threadIdx.x; idx < din * TP; idx += blockDim.x) {
                const int pt = idx / din, d = idx - pt * din;
                const long long p = p0 + pt;
                if (p < a.M) a.gin[p * din + d] = GIN[d * TP + pt];
            }
        }
        __syncthreads();
    }

    // flush the small-parameter accumulators into this CTA's partial-gradient vector
    for (int idx = threadIdx.x; idx < NS; idx += blockDim.x) {
        int dst;
        if (idx < sbL) {
            const int l = idx / W;
            dst = a.lay.off_b[l] + (idx - l * W);
        } else if (idx == sbL) {
            dst = a.lay.off_b[L];
        } else if (idx < swout) {
            dst = a.lay.off_w[0] + (idx - sw0);
        } else if (idx < sab) {
            dst = a.lay.off_w[L] + (idx - swout);
        } else {
            const int l = (idx - sab) / 7;
            dst = a.lay.off_ab[l] + (idx - sab - 7 * l);
        }
        gp[dst] += SACC[idx];
    }
}

inline size_t plain_bwd_smem_bytes(int W, int L, int din, int nbuf) {
    const int TP = 128, HS = TP + 4;
    const size_t ns = (((size_t)L * W + 1 + (size_t)W * din + W + 7 * L + 3) & ~(size_t)3) + HS;
    return sizeof(float) * ((size_t)2 * W * HS + (size_t)2 * din * TP + TP + ns +
                            (L > 1 ? (size_t)nbuf * W * plain_wp(W) : 0));
}
inline int plain_bwd_nbuf(int W, int L, int din) {
    const size_t s1 = plain_bwd_smem_bytes(W, L, din, 1), s2 = plain_bwd_smem_bytes(W, L, din, 2);
    const size_t sm_bytes = 227 * 1024;
    const int c1 = (int)(sm_bytes / (s1 + 1024)), c2 = (int)(sm_bytes / (s2 + 1024));
    return (c2 >= 1 && (c2 == c1 || c2 >= 3)) ? 2 : 1;
}

template <int ACT>
int plain_bwd_grid(int W, int L, int din, long long num_tiles, int* grid_out) {
    const int nw = (W + PG - 1) / PG;
    const size_t smem = plain_bwd_smem_bytes(W, L, din, plain_bwd_nbuf(W, L, din));
    const void* k = nw <= PLAIN_NARROW_WARPS ? (const void*)mlp_plain_bwd_kernel<ACT, true> : (const void*)mlp_plain_bwd_kernel<ACT, false>;
    return grid_for(k, nw * 32, smem, num_tiles, grid_out);
}

template <int ACT>
int launch_plain_bwd(const BwdArgs& a0, int grid, cudaStream_t stream) {
    BwdArgs a = a0;
    const int W = a.net.W, L = a.net.L, din = a.net.din;
    const int nw = (W + PG - 1) / PG;
    a.aux = plain_bwd_nbuf(W, L, din);
    const size_t smem = plain_bwd_smem_bytes(W, L, din, a.aux);
    if (nw <= PLAIN_NARROW_WARPS) mlp_plain_bwd_kernel<ACT, true><<<grid, nw * 32, smem, stream>>>(a);
    else                          mlp_plain_bwd_kernel<ACT, false><<<grid, nw * 32, smem, stream>>>(a);
    PDR_CUDA_CHECK(cudaGetLastError());
    return PDEREAD_OK;
}

}  // namespace pdr
