// Persistent MLP-jet kernels (sm_100a): forward of U / N with Taylor jets, and the matching
// reverse pass for the parameter gradients.
//
// Replaces, for every collocation point, the chain
//   Neural_Network.forward (reference Network.py:174-202)
//   -> nested torch.autograd.grad in Evaluate_Derivatives (PDE_Residual.py:92-145)
//   -> Loss.backward() (Test_Train.py:99-100)
// by propagating the x-/t-series of every neuron through the layers (jet_math.cuh) and by the
// hand-written adjoint of that propagation.
//
// A CTA owns a tile of TP = 32*PPL points and walks all layers of the network for it; a warp owns a chunk of TN = 10
// output neurons.  In the layer product
//   z_j[i] = sum_k W[i][k] h_j[k]      (j = jet index, i = output neuron)
// a thread keeps 10 (neuron, point) jets as accumulators in registers, reads the jets h[k] of its points from shared
// memory (rows [neuron][point][jet]) and its weights W[.][k] either through the read-only path or from a slice staged
// in shared memory by cp.async; which (neuron, point) pairs a thread owns depends on the jet length (ProdTile: the
// pair tile for J <= 4, the lane tile above).  The activation jets are applied to the accumulators in registers;
// activations never go to HBM except for the pre-activation stash the reverse pass needs (and, for networks too wide
// for shared memory, the SPILL instantiations that keep the tile's buffers in the workspace).
//
// Reverse pass, per layer (two shared buffers Zb / Yb that swap roles): weight + bias gradients as a
// register-tiled product of the two buffers; ybar = W^T zbar with the forward's tile and the activation
// VJP in its epilogue (ybar never leaves the registers); activations of the layer below from the stash.
#pragma once
#include <cuda_runtime.h>

#include <mutex>

#include "internal.h"
#include "jet_math.cuh"


namespace pdr {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n"
                 :
                 : "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

template <int ACT>
__device__ __forceinline__ void load_ab(const NetDev& net, int l, float* ab) {
    if (ACT == ACT_RATIONAL) {
#pragma unroll
        for (int t = 0; t < 4; ++t) ab[t] = __ldg(net.ra[l] + t);
#pragma unroll
        for (int t = 0; t < 3; ++t) ab[4 + t] = __ldg(net.rb[l] + t);
    } else {
#pragma unroll
        for (int t = 0; t < 7; ++t) ab[t] = 0.f;
    }
}


// Rows of the shared activation buffers keep the J jet coefficients of a point contiguous ([neuron][point][jet], row
// stride HS): one 16- / 8-byte access per point where J allows
template <int J>
__device__ __forceinline__ void row_load(const float* p, float* z) {
    if constexpr (J % 4 == 0) {
#pragma unroll
        for (int j = 0; j < J; j += 4) {
            const float4 v = *reinterpret_cast<const float4*>(p + j);
            z[j] = v.x; z[j + 1] = v.y; z[j + 2] = v.z; z[j + 3] = v.w;
        }
    } else if constexpr (J % 2 == 0) {
#pragma unroll
        for (int j = 0; j < J; j += 2) {
            const float2 v = *reinterpret_cast<const float2*>(p + j);
            z[j] = v.x; z[j + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < J; ++j) z[j] = p[j];
    }
}

// ---- register tiles of the layer products ----------------------------------------------------------------------
// Two thread mappings, chosen by the jet length (UsePairMap): both give a warp a chunk of TN = 10 output neurons and
// all points of the tile, both keep 10 (neuron, point) jets per thread and issue packed fma.rn.f32x2 (SASS FFMA2:
// two FMAs per instruction, same rounding as fmaf, one operand a broadcast scalar).
//   PairTile (J <= 4): lane = 2 a + b; b picks a half of the chunk (TH = 5 neurons), a a slot of points (a, a + 16, ..).
//     Per k-step: the 5 weights of the half (LDG.128 + LDG.32, two addresses per warp) and one 8- / 16-byte row
//     segment per point (16 distinct, contiguous); FFMA2 pairs are jet coefficients (2t, 2t+1) of a point, for odd J
//     the last coefficients of two points; the weight is the scalar operand.
//   LaneTile (J >= 5): lane = point, 10 neurons per thread.  Per k-step: three warp-uniform LDG.128 (10 weights as
//     pairs) and the J coefficients of the lane's point; FFMA2 pairs are neighbouring neurons, the activation is the
//     scalar operand.
// Measured on one B200, forward-U of the bench configurations (ms): J = 4 (C2, 100 k points) pair 0.265 / lane 0.320;
// J = 3 (C5, 10 M) 21.0 / 21.8; J = 5 (C3, 1 M) 4.15 / 3.71; J = 6 (C4, 524 k) 9.3 / 8.35 -- the long jets need more
// row loads per k-step in the pair tile (6 LDS.64 against 3) and lose.
template <int J>
struct UsePairMap {
    static constexpr bool v = (J <= PDR_PAIR_MAXJ);
};

// the J coefficients of one point from a row with point stride JP (p is 4 JP-byte aligned)
template <int J, int JP>
__device__ __forceinline__ void point_load(const float* p, float* z) {
    float t[JP];
    if constexpr (JP % 4 == 0) {
#pragma unroll
        for (int j = 0; j < JP; j += 4) {
            const float4 v = *reinterpret_cast<const float4*>(p + j);
            t[j] = v.x; t[j + 1] = v.y; t[j + 2] = v.z; t[j + 3] = v.w;
        }
    } else if constexpr (JP % 2 == 0) {
#pragma unroll
        for (int j = 0; j < JP; j += 2) {
            const float2 v = *reinterpret_cast<const float2*>(p + j);
            t[j] = v.x; t[j + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < JP; ++j) t[j] = p[j];
    }
#pragma unroll
    for (int j = 0; j < J; ++j) z[j] = t[j];
}
template <int J, int JP>
__device__ __forceinline__ void point_store(float* p, const float* z) {
    float t[JP];
#pragma unroll
    for (int j = 0; j < JP; ++j) t[j] = (j < J) ? z[j] : 0.f;
    if constexpr (JP % 4 == 0) {
#pragma unroll
        for (int j = 0; j < JP; j += 4) *reinterpret_cast<float4*>(p + j) = make_float4(t[j], t[j + 1], t[j + 2], t[j + 3]);
    } else if constexpr (JP % 2 == 0) {
#pragma unroll
        for (int j = 0; j < JP; j += 2) *reinterpret_cast<float2*>(p + j) = make_float2(t[j], t[j + 1]);
    } else {
#pragma unroll
        for (int j = 0; j < JP; ++j) p[j] = t[j];
    }
}

template <bool PAIR, int PPL, int J>
struct ProdTile;

template <int PPL, int J>
struct ProdTile<true, PPL, J> {
    static constexpr int NPT = TH;              // neurons per thread
    static constexpr int QP = 2 * PPL;          // points per thread
    static constexpr int WROW = TNP_PAIR;       // packed floats per chunk and k
    // point stride of the forward kernel's rows: odd jet lengths padded to an even stride (8- / 16-byte row accesses;
    // J = 3 at C5: 21.0 ms against 22.4 unpadded).  The reverse kernel keeps J: its weight-gradient tiles reduce over
    // whole rows.
    static constexpr int JP_FWD = (J == 1) ? 1 : ((J + 1) & ~1);
    static constexpr int JE = J / 2;
    float2 e[TH][QP][JE > 0 ? JE : 1];
    float2 o[TH][QP / 2];
    static __device__ __forceinline__ int neuron(int lane, int c, int n) { return c * TN + (lane & 1) * TH + n; }
    static __device__ __forceinline__ int point(int lane, int q) { return (lane >> 1) + 16 * q; }
    static __device__ __forceinline__ int wofs(int lane, int c) { return c * TNP_PAIR + (lane & 1) * THP; }
    __device__ __forceinline__ float get(int n, int q, int j) const {
        if (j < 2 * JE) return (j & 1) ? e[n][q][j >> 1].y : e[n][q][j >> 1].x;
        return (q & 1) ? o[n][q >> 1].y : o[n][q >> 1].x;
    }
    __device__ __forceinline__ void set(int n, int q, int j, float v) {
        if (j < 2 * JE) {
            if (j & 1) e[n][q][j >> 1].y = v; else e[n][q][j >> 1].x = v;
        } else {
            if (q & 1) o[n][q >> 1].y = v; else o[n][q >> 1].x = v;
        }
    }
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int n = 0; n < TH; ++n) {
#pragma unroll
            for (int q = 0; q < QP; ++q)
#pragma unroll
                for (int t = 0; t < JE; ++t) e[n][q][t] = make_float2(0.f, 0.f);
#pragma unroll
            for (int s = 0; s < QP / 2; ++s) o[n][s] = make_float2(0.f, 0.f);
        }
    }
    // one k-step: acc[n][q][j] += w[n] h[q][j].  hk: row k of the activation buffer at this thread's first point,
    // wk: this thread's weights of row k
    template <int JP, bool WS>
    __device__ __forceinline__ void step(const float* __restrict__ hk, const float* __restrict__ wk) {
        const float4 w4 = WS ? *reinterpret_cast<const float4*>(wk) : __ldg(reinterpret_cast<const float4*>(wk));
        const float w5 = WS ? wk[4] : __ldg(wk + 4);
        const float wv[TH] = {w4.x, w4.y, w4.z, w4.w, w5};
        float hv[QP][J];
#pragma unroll
        for (int q = 0; q < QP; ++q) point_load<J, JP>(hk + 16 * q * JP, hv[q]);
#pragma unroll
        for (int q = 0; q < QP; ++q)
#pragma unroll
            for (int t = 0; t < JE; ++t) {
                const float2 h2 = make_float2(hv[q][2 * t], hv[q][2 * t + 1]);
#pragma unroll
                for (int n = 0; n < TH; ++n) e[n][q][t] = __ffma2_rn(make_float2(wv[n], wv[n]), h2, e[n][q][t]);
            }
        if constexpr (J & 1) {
#pragma unroll
            for (int s = 0; s < QP / 2; ++s) {
                const float2 h2 = make_float2(hv[2 * s][J - 1], hv[2 * s + 1][J - 1]);
#pragma unroll
                for (int n = 0; n < TH; ++n) o[n][s] = __ffma2_rn(make_float2(wv[n], wv[n]), h2, o[n][s]);
            }
        }
    }
};

template <int PPL, int J>
struct ProdTile<false, PPL, J> {
    static constexpr int NPT = TN;
    static constexpr int QP = PPL;
    static constexpr int WROW = TNP_LANE;
    static constexpr int JP_FWD = J;
    float2 a[QP][J][TN / 2];
    static __device__ __forceinline__ int neuron(int, int c, int n) { return c * TN + n; }
    static __device__ __forceinline__ int point(int lane, int q) { return lane + 32 * q; }
    static __device__ __forceinline__ int wofs(int, int c) { return c * TNP_LANE; }
    __device__ __forceinline__ float get(int n, int q, int j) const { return (n & 1) ? a[q][j][n >> 1].y : a[q][j][n >> 1].x; }
    __device__ __forceinline__ void set(int n, int q, int j, float v) {
        if (n & 1) a[q][j][n >> 1].y = v; else a[q][j][n >> 1].x = v;
    }
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int q = 0; q < QP; ++q)
#pragma unroll
            for (int j = 0; j < J; ++j)
#pragma unroll
                for (int n = 0; n < TN / 2; ++n) a[q][j][n] = make_float2(0.f, 0.f);
    }
    template <int JP, bool WS>
    __device__ __forceinline__ void step(const float* __restrict__ hk, const float* __restrict__ wk) {
        const float4* w4p = reinterpret_cast<const float4*>(wk);
        const float4 b0 = WS ? w4p[0] : __ldg(w4p);
        const float4 b1 = WS ? w4p[1] : __ldg(w4p + 1);
        const float4 b2 = WS ? w4p[2] : __ldg(w4p + 2);
        const float2 bw[TNP_LANE / 2] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y),
                                         make_float2(b1.z, b1.w), make_float2(b2.x, b2.y), make_float2(b2.z, b2.w)};
#pragma unroll
        for (int q = 0; q < QP; ++q) {
            float hv[J];
            point_load<J, JP>(hk + q * 32 * JP, hv);
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const float2 a2 = make_float2(hv[j], hv[j]);
#pragma unroll
                for (int n = 0; n < TN / 2; ++n) a[q][j][n] = __ffma2_rn(a2, bw[n], a[q][j][n]);
            }
        }
    }
};

// The whole reduction of one layer product for this thread's register tile.  WS: the weights were staged in shared
// memory (plain loads), else they come through the read-only path.
template <int JP, bool WS, class Tile>
__device__ __forceinline__ void product_loop(Tile& acc, const float* __restrict__ h, int HS,
                                             const float* __restrict__ w, int wstride, int W) {
    constexpr int UNROLL = PDR_PROD_UNROLL;
#pragma unroll UNROLL
    for (int k = 0; k < W; ++k) {
        acc.template step<JP, WS>(h, w);
        w += wstride;
        h += HS;
    }
}

// Stash rows keep the J jet coefficients of a point contiguous ([.. neuron][point][jet]) so that one lane moves
// them with a single 16- / 8-byte access when J is a multiple of 4 / 2 (consecutive lanes -> consecutive addresses).
// The same helper stores rows of the shared activation buffers.
template <int J>
__device__ __forceinline__ void stash_store(float* p, const float* z) {
    if constexpr (J % 4 == 0) {
#pragma unroll
        for (int j = 0; j < J; j += 4) *reinterpret_cast<float4*>(p + j) = make_float4(z[j], z[j + 1], z[j + 2], z[j + 3]);
    } else if constexpr (J % 2 == 0) {
#pragma unroll
        for (int j = 0; j < J; j += 2) *reinterpret_cast<float2*>(p + j) = make_float2(z[j], z[j + 1]);
    } else {
#pragma unroll
        for (int j = 0; j < J; ++j) p[j] = z[j];
    }
}
template <int J>
__device__ __forceinline__ void stash_load(const float* p, float* z) {
    if constexpr (J % 4 == 0) {
#pragma unroll
        for (int j = 0; j < J; j += 4) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(p + j));
            z[j] = v.x; z[j + 1] = v.y; z[j + 2] = v.z; z[j + 3] = v.w;
        }
    } else if constexpr (J % 2 == 0) {
#pragma unroll
        for (int j = 0; j < J; j += 2) {
            const float2 v = __ldcs(reinterpret_cast<const float2*>(p + j));
            z[j] = v.x; z[j + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < J; ++j) z[j] = __ldcs(p + j);
    }
}

// pre-activation jet of hidden layer 0 (the input layer) for neuron i at tile point pt:
// z_0 = b + sum_d W[i][d] in[d];  the x-series is seeded by input column 1 and the t-series by
// input column 0 (coords are (t, x), reference PDE_Residual.py:33-34).
template <int NX, int NT>
__device__ __forceinline__ void input_layer_jet(const NetDev& net, const float* IN, int TP, int i, int pt,
                                                float* z) {
    constexpr int J = 1 + NX + NT;
    const float* wrow = net.w[0] + (size_t)i * net.din;
    float s = __ldg(net.b[0] + i);
    for (int d = 0; d < net.din; ++d) s = fmaf(__ldg(wrow + d), IN[d * TP + pt], s);
#pragma unroll
    for (int j = 0; j < J; ++j) z[j] = 0.f;
    z[0] = s;
    if (NX > 0) z[1] = __ldg(wrow + 1);
    if (NT > 0) z[NX + 1] = __ldg(wrow + 0);
}

// =============================================================================================
// forward
// =============================================================================================
// PPL = points per lane on average (tile = 32 PPL points): PointsPerLane<J> (the tile the reverse kernel and the
// stash use), or FWD_ONLY_PPL_J1 for forward-only calls of the plain MLP (J = 1).
// (Measured and dropped: a single activation buffer updated in place -- outputs wait in their accumulator registers
// for a barrier, then replace the layer below.  It halves the shared memory, but costs a barrier per layer and, where
// it buys a second CTA (width 100, J = 6), a 96-register cap: C4 forward 8.59 ms against 8.28, C3 4.18 against 3.72.)
// SPILL: the two activation buffers live in this CTA's slice of the workspace (a.hbuf) instead of shared memory --
// networks too wide for an SM's shared memory.  A separate instantiation, so that the common case keeps its
// shared-memory address space (LDS / STS instead of generic loads: reverse-U at C2 0.66 ms against 0.94).
// MODE (forward kernel): 0 = buffers in shared memory, weights through the read-only path; 1 = the same with per-warp
// weight slices staged in shared memory; 2 = SPILL.  Compile-time, not a run-time flag: ptxas' schedule of the product
// loop in mode 0 changes (C3 forward 3.71 -> 4.30 ms) as soon as the staging code shares its function.
constexpr int FWD_PLAIN = 0, FWD_WSTAGE = 1, FWD_SPILL = 2;
template <int ACT, int NX, int NT, bool STASH, int PPL, int MODE>
__global__ void __launch_bounds__(RegCap<1 + NX + NT>::max_threads) mlp_jet_fwd_kernel(const FwdArgs a) {
    constexpr bool SPILL = (MODE == FWD_SPILL);
    constexpr int J = 1 + NX + NT;
    using Tile = ProdTile<UsePairMap<J>::v, PPL, J>;
    constexpr int JP = Tile::JP_FWD;
    constexpr int TP = 32 * PPL;
    constexpr int QP = Tile::QP, NPT = Tile::NPT;
    constexpr int HS = JP * TP + 4;

    extern __shared__ __align__(16) float smem[];
    const NetDev& net = a.net;
    const int W = net.W, L = net.L, din = net.din, NC = net.NC;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;

    // activation buffers: shared memory, or this CTA's slice of the workspace when the network is too wide for it
    // (block barriers order the CTA's own global-memory accesses just as they order shared memory)
    // (keep these three expressions as they are: ptxas' schedule of the product loop below flips with the way they
    //  are written -- J = 6 forward 8.3 ms against 10.2 for `smem + 2 W HS`; check C4 after touching this kernel)
    float* Hbuf0 = smem;
    float* Hbuf1 = Hbuf0 + (size_t)W * HS;
    float* IN = Hbuf1 + (size_t)W * HS;
    if constexpr (SPILL) {
        Hbuf0 = a.hbuf + (size_t)blockIdx.x * 2 * W * HS;
        Hbuf1 = Hbuf0 + (size_t)W * HS;
        IN = smem;
    }
    float* RED = IN + din * TP;   // [nwarps][J][TP]

    double lsum = 0.0;
    const float* wout = net.w[L];

    // MODE 1 (one chunk per warp, room in shared memory): every warp keeps ITS slice of the layer's packed weights
    // -- the WROW columns of its chunk, all W rows -- in shared memory.  No other warp reads that slice, so the warp
    // requests the next layer's copy (cp.async) as soon as its own product loop is done and the copy overlaps its
    // activation epilogue; the barrier that ends the layer also publishes the copy.
    constexpr bool wstage = (MODE == FWD_WSTAGE);
    [[maybe_unused]] float* WSL = RED + (size_t)nwarps * J * TP + (size_t)warp * W * Tile::WROW;     // this warp's slice [W][WROW]
    [[maybe_unused]] auto stage_slice = [&](int l) {
        const float* src = a.wt_packed + (size_t)(l - 1) * W * NC * Tile::WROW + warp * Tile::WROW;
        constexpr int V = Tile::WROW / 4;
        for (int idx = lane; idx < W * V; idx += 32) {
            const int row = idx / V, part4 = idx - row * V;
            cp_async16(WSL + row * Tile::WROW + part4 * 4, src + (size_t)row * NC * Tile::WROW + part4 * 4);
        }
        cp_async_commit();
    };

    for (long long tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const long long p0 = tile * TP;
        if constexpr (wstage) {
            if (L > 1) stage_slice(1);             // (the slice was last read before the barrier that ended the previous tile)
        }
        for (int idx = threadIdx.x; idx < din * TP; idx += blockDim.x) {
            const int pt = idx / din, d = idx - pt * din;
            const long long p = p0 + pt;
            IN[d * TP + pt] = (p < a.M) ? __ldg(a.in + p * din + d) : 0.f;
        }
        __syncthreads();

        float part[QP][J];
#pragma unroll
        for (int q = 0; q < QP; ++q)
#pragma unroll
            for (int j = 0; j < J; ++j) part[q][j] = 0.f;

        float* Hin = Hbuf0;
        float* Hout = Hbuf1;
        float ab[7];

        // ---- hidden layer 0 -------------------------------------------------------------
        load_ab<ACT>(net, 0, ab);
        for (int c = warp; c < NC; c += nwarps) {
#pragma unroll 1
            for (int n = 0; n < NPT; ++n) {
                const int i = Tile::neuron(lane, c, n);
                if (i < W) {
                    const float wo = (L == 1) ? __ldg(wout + i) : 0.f;
#pragma unroll
                    for (int q = 0; q < QP; ++q) {
                        const int pt = Tile::point(lane, q);
                        float z[J], y[J];
                        input_layer_jet<NX, NT>(net, IN, TP, i, pt, z);
                        act_fwd<ACT, NX, NT, float>(z, ab, y);
                        if (L == 1) {
#pragma unroll
                            for (int j = 0; j < J; ++j) part[q][j] = fmaf(wo, y[j], part[q][j]);
                        } else {
                            point_store<J, JP>(Hin + (size_t)i * HS + pt * JP, y);
                        }
                    }
                }
            }
        }
        if constexpr (wstage) cp_async_wait_all();
        __syncthreads();

        // ---- hidden layers 1 .. L-1 ------------------------------------------------------
        for (int l = 1; l < L; ++l) {
            const float* wp = a.wt_packed + (size_t)(l - 1) * W * NC * Tile::WROW;
            const float* bias = net.b[l];
            const bool last = (l == L - 1);
            load_ab<ACT>(net, l, ab);
            for (int c = warp; c < NC; c += nwarps) {
                Tile acc;
                acc.zero();
#pragma unroll
                for (int n = 0; n < NPT; ++n) {
                    const int i = Tile::neuron(lane, c, n);
                    const float bv = (i < W) ? __ldg(bias + i) : 0.f;
#pragma unroll
                    for (int q = 0; q < QP; ++q) acc.set(n, q, 0, bv);
                }
                if constexpr (wstage) {
                    product_loop<JP, true>(acc, Hin + Tile::point(lane, 0) * JP, HS, WSL + Tile::wofs(lane, 0), Tile::WROW, W);
                    if (!last) stage_slice(l + 1);
                } else {
                    product_loop<JP, false>(acc, Hin + Tile::point(lane, 0) * JP, HS, wp + Tile::wofs(lane, c), NC * Tile::WROW, W);
                }
#pragma unroll
                for (int n = 0; n < NPT; ++n) {
                    const int i = Tile::neuron(lane, c, n);
                    if (i < W) {
                        const float wo = last ? __ldg(wout + i) : 0.f;
#pragma unroll
                        for (int q = 0; q < QP; ++q) {
                            const int pt = Tile::point(lane, q);
                            float z[J], y[J];
#pragma unroll
                            for (int j = 0; j < J; ++j) z[j] = acc.get(n, q, j);
                            act_fwd<ACT, NX, NT, float>(z, ab, y);
                            if (STASH) {
                                if (ACT == ACT_TANH) z[0] = y[0];       // see "rows of the reverse pass"
                                stash_store<J>(a.stash + (((size_t)tile * (L - 1) + (l - 1)) * W + i) * (J * TP) + pt * J, z);
                            }
                            if (last) {
#pragma unroll
                                for (int j = 0; j < J; ++j) part[q][j] = fmaf(wo, y[j], part[q][j]);
                            } else {
                                point_store<J, JP>(Hout + (size_t)i * HS + pt * JP, y);
                            }
                        }
                    }
                }
            }
            if constexpr (wstage) cp_async_wait_all();
            __syncthreads();
            float* t = Hin; Hin = Hout; Hout = t;
        }

        // ---- output layer: (pair tile: the two halves of a chunk meet in a lane pair,) the chunks in shared memory
#pragma unroll
        for (int q = 0; q < QP; ++q)
#pragma unroll
            for (int j = 0; j < J; ++j) {
                float v = part[q][j];
                if constexpr (UsePairMap<J>::v) {
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    if ((lane & 1) == 0) RED[(warp * J + j) * TP + Tile::point(lane, q)] = v;
                } else {
                    RED[(warp * J + j) * TP + Tile::point(lane, q)] = v;
                }
            }
        __syncthreads();
        const float bout = __ldg(net.b[L]);
        for (int idx = threadIdx.x; idx < J * TP; idx += blockDim.x) {
            const int j = idx / TP, pt = idx - j * TP;
            const long long p = p0 + pt;
            float u = (j == 0) ? bout : 0.f;
            for (int w = 0; w < nwarps; ++w) u += RED[(w * J + j) * TP + pt];
            if (p >= a.M) continue;
            if (J == 1) {
                if (a.dtm_in != nullptr) {
                    const float R = __ldg(a.dtm_in + p) - u;
                    if (a.resid_out) a.resid_out[p] = R;
                    if (a.seed_out)
                        a.seed_out[p] = (a.grad_resid != nullptr) ? -__ldg(a.grad_resid + p) : -a.seed_scale * R;
                    lsum += (double)R * (double)R;
                } else {
                    a.out0[p] = u;
                }
            } else {
                if (j <= NX) {
                    float f = 1.f;
#pragma unroll
                    for (int t = 2; t <= NX; ++t) if (t <= j) f *= (float)t;
                    a.out0[p * (NX + 1) + j] = u * f;
                } else if (j == NX + NT && a.out1 != nullptr) {
                    float f = 1.f;
#pragma unroll
                    for (int t = 2; t <= NT; ++t) f *= (float)t;
                    a.out1[p] = u * f;
                }
            }
        }
        // RED / IN / H are next written only after further barriers of the next tile
        __syncthreads();
    }

    if (J == 1 && a.sumsq != nullptr) {
        // block reduction of the squared residuals (double), one atomic per CTA
        __shared__ double red_d[MAX_WARPS];
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

// =============================================================================================
// backward
// =============================================================================================

// weight-gradient micro kernel: gW[i][k] += sum_r ZB[i][r] * HP[k][r],  r over (jet, point).
// The phase is bound by shared-memory wavefronts, not FMAs, so the lane -> output mapping is chosen for the
// cheapest 16-byte loads the hardware offers (measured with tests/lds_probe: a warp-wide LDS.128 costs 2
// wavefronts when the lanes' addresses form runs of equal addresses over consecutive rows (row stride = 4 mod 32
// banks) or alternate between two rows, 4 or more for modulo / scattered patterns):
//   lane = 2 a + b;  rows i = i0 + b + 2 ti (2 rows per load, alternating lanes),
//                    columns k = k0 + a + 16 tk (16 consecutive rows per load, runs of 2 lanes)
// so that every load is a 2-wavefront one and a warp owns a (2 TI) x (16 TK) block of the gradient; TI = 5 makes
// the row block one neuron chunk (TN = 10 rows), so a block of NC warps covers a layer in one round per column block.
// Column k = W is the bias gradient: its "activation" row is ONES (1 over the order-0 block of the row, 0 over the
// higher jet coefficients), so gb[i] = sum_p ZB[i][0][p] falls out of the same tiles.
template <int TI, int TK>
__device__ __forceinline__ void wgrad_tiles(const float* __restrict__ ZB, const float* __restrict__ HP,
                                            const float* __restrict__ ONES, int W, int HS, int R,
                                            float* __restrict__ gw, float* __restrict__ gb) {
    constexpr int RT = 2 * TI, CT = 16 * TK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int la = lane >> 1, lb = lane & 1;
    const int WK = W + 1;   // + the bias column
    const int nrt = (W + RT - 1) / RT, nct = (WK + CT - 1) / CT;
    for (int task = warp; task < nrt * nct; task += nwarps) {
        const int rt = task / nct, ct = task - rt * nct;
        const int i0 = rt * RT + lb, k0 = ct * CT + la;
        const float4* ap[TI];
        const float4* bp[TK];
#pragma unroll
        for (int ti = 0; ti < TI; ++ti) {
            const int i = i0 + 2 * ti;
            ap[ti] = reinterpret_cast<const float4*>(ZB + (size_t)(i < W ? i : W - 1) * HS);
        }
#pragma unroll
        for (int tk = 0; tk < TK; ++tk) {
            const int k = k0 + 16 * tk;
            bp[tk] = reinterpret_cast<const float4*>(
                k < W ? HP + (size_t)k * HS : (k == W ? ONES : HP + (size_t)(W - 1) * HS));
        }
        // packed fma.rn.f32x2: each output keeps two partial sums (even / odd reduction index), the operand
        // pairs are the .xy / .zw halves of the 16-byte loads
        float2 acc2[TI][TK];
#pragma unroll
        for (int ti = 0; ti < TI; ++ti)
#pragma unroll
            for (int tk = 0; tk < TK; ++tk) acc2[ti][tk] = make_float2(0.f, 0.f);
        const int R4 = R >> 2;
#pragma unroll 1
        for (int r = 0; r < R4; ++r) {
            float4 av[TI];
#pragma unroll
            for (int ti = 0; ti < TI; ++ti) av[ti] = ap[ti][r];
#pragma unroll
            for (int tk = 0; tk < TK; ++tk) {
                const float4 bv = bp[tk][r];
#pragma unroll
                for (int ti = 0; ti < TI; ++ti) {
                    float2 s2 = acc2[ti][tk];
                    s2 = __ffma2_rn(make_float2(av[ti].x, av[ti].y), make_float2(bv.x, bv.y), s2);
                    s2 = __ffma2_rn(make_float2(av[ti].z, av[ti].w), make_float2(bv.z, bv.w), s2);
                    acc2[ti][tk] = s2;
                }
            }
        }
#pragma unroll
        for (int ti = 0; ti < TI; ++ti) {
            const int i = i0 + 2 * ti;
            if (i < W) {
#pragma unroll
                for (int tk = 0; tk < TK; ++tk) {
                    const int k = k0 + 16 * tk;
                    const float v = acc2[ti][tk].x + acc2[ti][tk].y;
                    if (k < W) atomicAdd(&gw[(size_t)i * W + k], v);   // private to this CTA; RED avoids the load round trip
                    else if (k == W) atomicAdd(&gb[i], v);                 // shared-memory accumulator
                }
            }
        }
    }
}

// ---- rows of the reverse pass ------------------------------------------------------------------------------
// A "row" is the pre-activation jet of one (neuron, point) as the reverse kernel sees it.  For tanh its slot 0
// holds y_0 = tanh(z_0) instead of z_0 (written so by the forward kernel's stash): every other coefficient of the
// tanh jet and of its adjoint depends on z_0 only through y_0, so the reverse pass never evaluates tanh, and the
// VJP takes the forward jet y from shared memory (it is there for the weight gradient) instead of recomputing it.
// The rational VJP reuses y the same way (its numerator and series division are skipped; z itself stays in the row).
template <int ACT>
struct VJP_FROM_Y {
    static constexpr bool v = (ACT == ACT_TANH || ACT == ACT_RATIONAL);
};

template <int ACT, int NX, int NT>
__device__ __forceinline__ void input_layer_row(const NetDev& net, const float* IN, int TP, int i, int pt, float* z) {
    input_layer_jet<NX, NT>(net, IN, TP, i, pt, z);
    if (ACT == ACT_TANH) z[0] = pdr_tanh(z[0]);
}

template <int ACT, int NX, int NT>
__device__ __forceinline__ void act_fwd_row(const float* z, const float* ab, float* y) {
    if (ACT == ACT_TANH) {
        float w[1 + NX + NT];
        tanh_series<NX, NT, float>(z[0], z, y, w);
    } else {
        act_fwd<ACT, NX, NT, float>(z, ab, y);
    }
}

template <int ACT, int NX, int NT>
__device__ __forceinline__ void act_vjp_row(const float* z, const float* y, const float* ab, const float* ybar,
                                            float* zbar, float* abbar) {
    if (ACT == ACT_TANH)          tanh_vjp_y<NX, NT, float>(z, y, ybar, zbar);
    else if (ACT == ACT_RATIONAL) rational_vjp_y<NX, NT, float>(z, y, ab, ybar, zbar, abbar);
    else                          act_vjp<ACT, NX, NT, float>(z, ab, ybar, zbar, abbar);
}

// FUSED: one warp per neuron chunk (blockDim = 32 NC), the VJP runs in the product epilogue.  !FUSED ("wide"
// blocks of BWD_WIDE_WARPS > NC warps, one CTA per SM): the VJP runs in phase (d), where all warps take part.
template <int ACT, int NX, int NT, bool FUSED, bool SPILL>
__global__ void __launch_bounds__(RegCap<1 + NX + NT>::max_threads) mlp_jet_bwd_kernel(const BwdArgs a) {
    constexpr int J = 1 + NX + NT;
    constexpr int PPL = PointsPerLane<J>::v;
    constexpr int TP = 32 * PPL;
    using Tile = ProdTile<UsePairMap<J>::v, PPL, J>;
    constexpr int QP = Tile::QP, NPT = Tile::NPT;
    constexpr int HS = J * TP + 4;

    extern __shared__ __align__(16) float smem[];
    const NetDev& net = a.net;
    const int W = net.W, L = net.L, din = net.din, NC = net.NC;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;

    // (a.hbuf: the two buffers live in this CTA's slice of the workspace -- networks too wide for shared memory)
    float* Zb = smem;                          // zbar of the layer above the current weight matrix ...
    float* Yb = Zb + (size_t)W * HS;           // ... activations of the layer below it (the two swap every layer)
    float* IN = Yb + (size_t)W * HS;           // [din][TP]
    if constexpr (SPILL) {
        Zb = a.hbuf + (size_t)blockIdx.x * 2 * W * HS;
        Yb = Zb + (size_t)W * HS;
        IN = smem;
    }
    float* SD = IN + din * TP;                 // [J][TP] seeds = d(loss)/d(u_j), factorials folded in
    float* GIN = SD + J * TP;                  // [din][TP]
    float* SACC = GIN + din * TP;              // small-parameter gradient accumulators (persist over tiles)
    const int sbL = L * W;                     // output bias
    const int sw0 = sbL + 1;                   // input-layer weights [W][din]
    const int swout = sw0 + W * din;           // output weights [W]
    const int sab = swout + W;                 // rational coefficients [L][7]
    const int NS = sab + 7 * L;
    float* ONES = SACC + ((NS + 3) & ~3);      // [HS] 1 at the order-0 coefficient of every point, 0 elsewhere (bias column)
    // a.aux != 0: the weights of the matrix below are copied into shared memory (cp.async) while the weight-gradient
    // tiles of phase (b) run, and the product of phase (c) reads them from there
    float* WSB = ONES + ((HS + 3) & ~3);       // [W][NC * WROW]
    const bool stage = a.aux != 0;

    // activation-only phases: one work item = G neurons; blocks with more warps than neuron chunks (wide
    // networks, one CTA per SM) hand out single neurons so that all warps take part
    constexpr bool fused = FUSED;
    const int G = fused ? TN : 1;
    const int nitems = (W + G - 1) / G;

    float* gp = a.gpart + (size_t)blockIdx.x * a.lay.total;
    if (!a.accumulate)
        for (int idx = threadIdx.x; idx < a.lay.total; idx += blockDim.x) gp[idx] = 0.f;
    for (int idx = threadIdx.x; idx < NS; idx += blockDim.x) SACC[idx] = 0.f;
    for (int idx = threadIdx.x; idx < HS; idx += blockDim.x) ONES[idx] = (idx < J * TP && idx % J == 0) ? 1.f : 0.f;
    __syncthreads();

    const float* wout = net.w[L];

    for (long long tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const long long p0 = tile * TP;
        for (int idx = threadIdx.x; idx < din * TP; idx += blockDim.x) {
            const int pt = idx / din, d = idx - pt * din;
            const long long p = p0 + pt;
            IN[d * TP + pt] = (p < a.M) ? __ldg(a.in + p * din + d) : 0.f;
            GIN[d * TP + pt] = 0.f;
        }
        for (int idx = threadIdx.x; idx < J * TP; idx += blockDim.x) {
            const int j = idx / TP, pt = idx - j * TP;
            const long long p = p0 + pt;
            float s = 0.f;
            if (p < a.M) {
                if (J == 1) {
                    s = a.g0 ? a.g0_scale * __ldg(a.g0 + p) : 0.f;
                } else if (j <= NX) {
                    float f = 1.f;
#pragma unroll
                    for (int t = 2; t <= NX; ++t) if (t <= j) f *= (float)t;
                    s = a.g0 ? a.g0_scale * f * __ldg(a.g0 + p * (NX + 1) + j) : 0.f;
                } else if (j == NX + NT) {
                    float f = 1.f;
#pragma unroll
                    for (int t = 2; t <= NT; ++t) f *= (float)t;
                    s = a.g1 ? a.g1_scale * f * __ldg(a.g1 + p) : 0.f;
                }
            }
            SD[idx] = s;
        }
        __syncthreads();

        // output bias: d/d b_out = sum_p seed_0
        if (warp == 0) {
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < PPL; ++q) s += SD[q * 32 + lane];
            s = warp_sum(s);
            if (lane == 0) SACC[sbL] += s;
        }

        // ---- top hidden layer: ybar = w_out * seed.  Zb <- zbar^{L-1}, Yb <- y^{L-2} (for the first weight gradient)
        {
            const int lv = L - 1;
            float ab[7], ab_lo[7], abbar[7];
            load_ab<ACT>(net, lv, ab);
            load_ab<ACT>(net, lv > 0 ? lv - 1 : 0, ab_lo);
#pragma unroll
            for (int t = 0; t < 7; ++t) abbar[t] = 0.f;
            const float* st_z = a.stash + ((size_t)tile * (L - 1) + (lv > 0 ? lv - 1 : 0)) * W * (J * TP);
            const float* st_zl = a.stash + ((size_t)tile * (L - 1) + (lv > 1 ? lv - 2 : 0)) * W * (J * TP);
            for (int c = warp; c < nitems; c += nwarps) {
                const int iend = min((c + 1) * G, W);
#pragma unroll 1
                for (int i = c * G; i < iend; ++i) {
                    const float wo = __ldg(wout + i);
                    float gwo = 0.f;
#pragma unroll
                    for (int q = 0; q < PPL; ++q) {
                        const int pt = q * 32 + lane;
                        float z[J], y[J], ybar[J], zbar[J];
                        if (lv == 0) input_layer_row<ACT, NX, NT>(net, IN, TP, i, pt, z);
                        else         stash_load<J>(st_z + (size_t)i * (J * TP) + pt * J, z);
                        if (lv > 0) {
                            float zl[J], yl[J];
                            if (lv == 1) input_layer_row<ACT, NX, NT>(net, IN, TP, i, pt, zl);
                            else         stash_load<J>(st_zl + (size_t)i * (J * TP) + pt * J, zl);
                            act_fwd_row<ACT, NX, NT>(zl, ab_lo, yl);
                            stash_store<J>(Yb + (size_t)i * HS + pt * J, yl);
                        }
                        act_fwd_row<ACT, NX, NT>(z, ab, y);
#pragma unroll
                        for (int j = 0; j < J; ++j) {
                            const float sd = SD[j * TP + pt];
                            ybar[j] = wo * sd;
                            gwo = fmaf(sd, y[j], gwo);
                        }
                        act_vjp_row<ACT, NX, NT>(z, y, ab, ybar, zbar, abbar);
                        stash_store<J>(Zb + (size_t)i * HS + pt * J, zbar);
                    }
                    gwo = warp_sum(gwo);
                    if (lane == 0) SACC[swout + i] += gwo;
                }
            }
            if (ACT == ACT_RATIONAL) {
#pragma unroll
                for (int t = 0; t < 7; ++t) {
                    const float s = warp_sum(abbar[t]);
                    if (lane == 0) atomicAdd(&SACC[sab + 7 * lv + t], s);
                }
            }
            __syncthreads();
        }

        // ---- hidden layers below, top down.  At the head of a trip Zb = zbar^{lv+1}, Yb = y^{lv}.
        for (int lv = L - 2; lv >= 0; --lv) {
            const int l = lv + 1;   // weight matrix between hidden layers lv and lv+1
            if (stage) {
                const float4* src = reinterpret_cast<const float4*>(a.wn_packed + (size_t)(l - 1) * W * NC * Tile::WROW);
                float4* dst = reinterpret_cast<float4*>(WSB);
                for (int i = threadIdx.x; i < W * NC * Tile::WROW / 4; i += blockDim.x) cp_async16(dst + i, src + i);
                cp_async_commit();
            }
            // (b) weight gradient of matrix l:  gW[i][k] += sum_r Zb[i][r] Yb[k][r]  (+ bias column)
            {
                float* gw = gp + a.lay.off_w[l];
                if (a.wg_mode == 2)      wgrad_tiles<5, 4>(Zb, Yb, ONES, W, HS, J * TP, gw, SACC + l * W);
                else if (a.wg_mode == 1) wgrad_tiles<5, 3>(Zb, Yb, ONES, W, HS, J * TP, gw, SACC + l * W);
                else                     wgrad_tiles<5, 2>(Zb, Yb, ONES, W, HS, J * TP, gw, SACC + l * W);
            }
            if (stage) cp_async_wait_all();
            __syncthreads();
            // (c) ybar^{lv}[k] = sum_i W_l[i][k] Zb[i].  Fused blocks (one warp per neuron chunk): ybar stays in the
            //     registers of the thread that owns (k, point) and the activation VJP runs in the same thread,
            //     unrolled over the TN neurons of the chunk; zbar^{lv} replaces y^{lv} in Yb (elements private to the
            //     thread), Zb is only read.  Wide blocks (more warps than chunks) park ybar in Yb and leave the VJP
            //     to phase (d), where all warps take part.
            float ab[7], abbar[7];
            load_ab<ACT>(net, lv, ab);
#pragma unroll
            for (int t = 0; t < 7; ++t) abbar[t] = 0.f;
            const float* st_z = a.stash + ((size_t)tile * (L - 1) + (lv > 0 ? lv - 1 : 0)) * W * (J * TP);
            const float* wp = a.wn_packed + (size_t)(l - 1) * W * NC * Tile::WROW;
            for (int c = warp; c < NC; c += nwarps) {
                Tile acc;
                acc.zero();
                if (stage) product_loop<J, true>(acc, Zb + Tile::point(lane, 0) * J, HS, WSB + Tile::wofs(lane, c), NC * Tile::WROW, W);
                // (measured without effect or with a loss, see DESIGN.md section 4: weight rows requested two k-steps ahead
                //  in registers, L1::evict_last on the weight loads, streaming stores for the forward's stash)
                else       product_loop<J, false>(acc, Zb + Tile::point(lane, 0) * J, HS, wp + Tile::wofs(lane, c), NC * Tile::WROW, W);
                // (requesting the stash rows of the VJP one neuron ahead, the first ones before the product loop, was
                //  measured: reverse-U at C2 0.697 ms against 0.662 -- the 8 extra live registers cost more than the latency)
                if (fused) {
#pragma unroll
                    for (int n = 0; n < NPT; ++n) {
                        const int k = Tile::neuron(lane, c, n);
                        if (k < W) {
#pragma unroll
                            for (int q = 0; q < QP; ++q) {
                                const int pt = Tile::point(lane, q);
                                float z[J], y[J], ybar[J], zbar[J];
                                if (lv == 0) input_layer_row<ACT, NX, NT>(net, IN, TP, k, pt, z);
                                else         stash_load<J>(st_z + (size_t)k * (J * TP) + pt * J, z);
#pragma unroll
                                for (int j = 0; j < J; ++j) {
                                    ybar[j] = acc.get(n, q, j);
                                    y[j] = 0.f;
                                }
                                if (VJP_FROM_Y<ACT>::v) row_load<J>(Yb + (size_t)k * HS + pt * J, y);
                                act_vjp_row<ACT, NX, NT>(z, y, ab, ybar, zbar, abbar);
                                stash_store<J>(Yb + (size_t)k * HS + pt * J, zbar);
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int n = 0; n < NPT; ++n) {
                        const int k = Tile::neuron(lane, c, n);
                        if (k < W) {
#pragma unroll
                            for (int q = 0; q < QP; ++q) {
                                float ybar[J];
#pragma unroll
                                for (int j = 0; j < J; ++j) ybar[j] = acc.get(n, q, j);
                                stash_store<J>(Yb + (size_t)k * HS + Tile::point(lane, q) * J, ybar);
                            }
                        }
                    }
                }
            }
            if (fused && ACT == ACT_RATIONAL) {
#pragma unroll
                for (int t = 0; t < 7; ++t) {
                    const float s = warp_sum(abbar[t]);
                    if (lane == 0) atomicAdd(&SACC[sab + 7 * lv + t], s);
                }
            }
            __syncthreads();
            // (d) Zb (dead now) <- y^{lv-1} = act(z^{lv-1}) for the next weight gradient; wide blocks also run the
            //     VJP here: Yb <- zbar^{lv} from the ybar parked there
            if (lv > 0 || !fused) {
                float ab_lo[7];
                load_ab<ACT>(net, lv > 0 ? lv - 1 : 0, ab_lo);
                const float* st_zl = a.stash + ((size_t)tile * (L - 1) + (lv > 1 ? lv - 2 : 0)) * W * (J * TP);
                if (fused) {
                    for (int c = warp; c < nitems; c += nwarps) {
                        const int iend = min((c + 1) * TN, W);
#pragma unroll 2
                        for (int i = c * TN; i < iend; ++i) {
#pragma unroll
                            for (int q = 0; q < PPL; ++q) {
                                const int pt = q * 32 + lane;
                                float zl[J], yl[J];
                                if (lv == 1) input_layer_row<ACT, NX, NT>(net, IN, TP, i, pt, zl);
                                else         stash_load<J>(st_zl + (size_t)i * (J * TP) + pt * J, zl);
                                act_fwd_row<ACT, NX, NT>(zl, ab_lo, yl);
                                stash_store<J>(Zb + (size_t)i * HS + pt * J, yl);
                            }
                        }
                    }
                } else {
#pragma unroll 1
                    for (int i = warp; i < W; i += nwarps) {
#pragma unroll
                        for (int q = 0; q < PPL; ++q) {
                            const int pt = q * 32 + lane;
                            float z[J], y[J], ybar[J], zbar[J];
                            if (lv == 0) input_layer_row<ACT, NX, NT>(net, IN, TP, i, pt, z);
                            else         stash_load<J>(st_z + (size_t)i * (J * TP) + pt * J, z);
                            row_load<J>(Yb + (size_t)i * HS + pt * J, ybar);
                            if (lv > 0) {
                                float zl[J], yl[J];
                                if (lv == 1) input_layer_row<ACT, NX, NT>(net, IN, TP, i, pt, zl);
                                else         stash_load<J>(st_zl + (size_t)i * (J * TP) + pt * J, zl);
                                act_fwd_row<ACT, NX, NT>(zl, ab_lo, yl);
                                stash_store<J>(Zb + (size_t)i * HS + pt * J, yl);
                            }
                            // y was overwritten by the parked ybar: tanh rebuilds it from the row (cheap series),
                            // the others take the plain VJP, which evaluates the forward parts once
                            if (ACT == ACT_TANH) {
                                act_fwd_row<ACT, NX, NT>(z, ab, y);
                                act_vjp_row<ACT, NX, NT>(z, y, ab, ybar, zbar, abbar);
                            } else {
                                act_vjp<ACT, NX, NT, float>(z, ab, ybar, zbar, abbar);
                            }
                            stash_store<J>(Yb + (size_t)i * HS + pt * J, zbar);
                        }
                    }
                }
                if (!fused && ACT == ACT_RATIONAL) {
#pragma unroll
                    for (int t = 0; t < 7; ++t) {
                        const float s = warp_sum(abbar[t]);
                        if (lane == 0) atomicAdd(&SACC[sab + 7 * lv + t], s);
                    }
                }
                __syncthreads();
            }
            float* t = Zb; Zb = Yb; Yb = t;
        }

        // ---- input layer: Zb = zbar^0.  d/db_0[i] = sum_p zbar_0, d/dW_0[i][d] = sum_p zbar_0 in_d (+ the series seeds: the x-series enters
        //      through input column 1, the t-series through column 0), d/d in_d = sum_i W_0[i][d] zbar_0
        for (int c = warp; c < nitems; c += nwarps) {
            const int iend = min((c + 1) * G, W);
            float gi[PPL][MAXD];
#pragma unroll
            for (int q = 0; q < PPL; ++q)
#pragma unroll
                for (int d = 0; d < MAXD; ++d) gi[q][d] = 0.f;
#pragma unroll 1
            for (int i = c * G; i < iend; ++i) {
                const float* wrow = net.w[0] + (size_t)i * din;
                float gwin[MAXD], gb0 = 0.f;
#pragma unroll
                for (int d = 0; d < MAXD; ++d) gwin[d] = 0.f;
#pragma unroll
                for (int q = 0; q < PPL; ++q) {
                    const int pt = q * 32 + lane;
                    const float zb0 = Zb[(size_t)i * HS + pt * J];
                    gb0 += zb0;
                    if (NX > 0) gwin[1] += Zb[(size_t)i * HS + pt * J + 1];
                    if (NT > 0) gwin[0] += Zb[(size_t)i * HS + pt * J + NX + 1];
#pragma unroll
                    for (int d = 0; d < MAXD; ++d) {
                        if (d < din) {
                            gwin[d] = fmaf(zb0, IN[d * TP + pt], gwin[d]);
                            gi[q][d] = fmaf(__ldg(wrow + d), zb0, gi[q][d]);
                        }
                    }
                }
#pragma unroll
                for (int d = 0; d < MAXD; ++d)
                    if (d < din) gwin[d] = warp_sum(gwin[d]);
                gb0 = warp_sum(gb0);
                if (lane == 0) {
                    SACC[i] += gb0;
#pragma unroll
                    for (int d = 0; d < MAXD; ++d)
                        if (d < din) SACC[sw0 + i * din + d] += gwin[d];
                }
            }
            if (a.gin != nullptr) {
#pragma unroll
                for (int q = 0; q < PPL; ++q)
#pragma unroll
                    for (int d = 0; d < MAXD; ++d)
                        if (d < din) atomicAdd(&GIN[d * TP + q * 32 + lane], gi[q][d]);
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

// =============================================================================================
// launchers
// =============================================================================================
// Resident CTAs for (kernel, block, shared memory) on the current device: attribute + occupancy queries cost
// several microseconds each, a step launches four such kernels, so the answer is remembered.
// one_per_sm: the activation buffers are spilled to the workspace, which is sized for one CTA per SM.
inline int grid_for(const void* kernel, int threads, size_t smem, long long num_tiles, int* grid_out,
                    bool one_per_sm = false) {
    struct Entry { const void* k; int threads; size_t smem; int dev; long long resident; };
    constexpr int NCACHE = 32;
    // process-wide, like the attribute: autograd runs the reverse calls on its own thread
    static Entry cache[NCACHE];
    static int used = 0, next_slot = 0;
    static std::mutex mu;
    int dev = 0;
    PDR_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    long long resident = -1;
    for (int i = 0; i < used; ++i)
        if (cache[i].k == kernel && cache[i].threads == threads && cache[i].smem == smem && cache[i].dev == dev) {
            resident = cache[i].resident;
            break;
        }
    if (resident < 0) {
        int sms = 0, occ = 0;
        PDR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        int rc = ensure_dynamic_smem(kernel, smem);
        if (rc != PDEREAD_OK) return rc;
        PDR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
        if (occ < 1) return PDEREAD_ERR_UNSUPPORTED_NET;
        resident = (long long)sms * occ;
        const int slot = used < NCACHE ? used++ : (next_slot++ % NCACHE);
        cache[slot] = Entry{kernel, threads, smem, dev, resident};
    }
    long long g = resident;
    if (one_per_sm) {
        int sms = 0;
        PDR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (g > sms) g = sms;
    }
    if (g > num_tiles) g = num_tiles;
    if (g < 1) g = 1;
    *grid_out = (int)g;
    return PDEREAD_OK;
}

template <int ACT, int NX, int NT, bool STASH, int PPL>
int launch_fwd_tile(FwdArgs a, cudaStream_t stream) {
    constexpr int J = 1 + NX + NT;
    const int nw = block_warps(a.net.W);
    // (a.hbuf is only ever set for the default tile, PPL = PointsPerLane<J>)
    size_t smem = a.hbuf ? fwd_smem_small(a.net.W, a.net.din, J, nw) : fwd_smem_bytes_tile(a.net.W, a.net.din, J, nw, 32 * PPL);
    a.aux = 0;
    if (!a.hbuf) {
        const size_t extra = fwd_wstage_bytes(a.net.W, a.net.L, J, smem);
        if (extra) { smem += extra; a.aux = 1; }
    }
    a.num_tiles = (a.M + 32 * PPL - 1) / (32 * PPL);
    const int mode = a.hbuf ? FWD_SPILL : (a.aux ? FWD_WSTAGE : FWD_PLAIN);
    const void* k = mode == FWD_SPILL    ? (const void*)mlp_jet_fwd_kernel<ACT, NX, NT, STASH, PPL, FWD_SPILL>
                    : mode == FWD_WSTAGE ? (const void*)mlp_jet_fwd_kernel<ACT, NX, NT, STASH, PPL, FWD_WSTAGE>
                                         : (const void*)mlp_jet_fwd_kernel<ACT, NX, NT, STASH, PPL, FWD_PLAIN>;
    int grid = 0;
    int rc = grid_for(k, nw * 32, smem, a.num_tiles, &grid, a.hbuf != nullptr);
    if (rc != PDEREAD_OK) return rc;
    if (mode == FWD_SPILL)       mlp_jet_fwd_kernel<ACT, NX, NT, STASH, PPL, FWD_SPILL><<<grid, nw * 32, smem, stream>>>(a);
    else if (mode == FWD_WSTAGE) mlp_jet_fwd_kernel<ACT, NX, NT, STASH, PPL, FWD_WSTAGE><<<grid, nw * 32, smem, stream>>>(a);
    else                         mlp_jet_fwd_kernel<ACT, NX, NT, STASH, PPL, FWD_PLAIN><<<grid, nw * 32, smem, stream>>>(a);
    PDR_CUDA_CHECK(cudaGetLastError());
    return PDEREAD_OK;
}

template <int ACT, int NX, int NT>
int launch_fwd(const FwdArgs& a, cudaStream_t stream) {
    constexpr int J = 1 + NX + NT;
    constexpr int PPL = PointsPerLane<J>::v;
    if (a.stash) return launch_fwd_tile<ACT, NX, NT, true, PPL>(a, stream);
    if constexpr (J == 1 && FWD_ONLY_PPL_J1 != PPL) {
        // the larger tile only while two CTAs of it still fit an SM
        if (fwd_smem_bytes_tile(a.net.W, a.net.din, J, block_warps(a.net.W), 32 * FWD_ONLY_PPL_J1) <= 112 * 1024)
            return launch_fwd_tile<ACT, NX, NT, false, FWD_ONLY_PPL_J1>(a, stream);
    }
    return launch_fwd_tile<ACT, NX, NT, false, PPL>(a, stream);
}

// The backward launch needs the grid size before the call (gpart is sized by it), so the
// occupancy query and the launch are separate.
template <int ACT, int NX, int NT>
int bwd_grid(int W, int L, int din, long long num_tiles, int* grid_out) {
    constexpr int J = 1 + NX + NT;
    const int nw = bwd_block_warps(W, L, din, J);
    const bool spill = bwd_spills(W, L, din, J);
    const size_t smem = spill ? bwd_smem_small(W, L, din, J) : bwd_smem_bytes(W, L, din, J);
    const int NC = (W + TN - 1) / TN;
    const void* k = spill ? ((nw <= NC) ? (const void*)mlp_jet_bwd_kernel<ACT, NX, NT, true, true>
                                        : (const void*)mlp_jet_bwd_kernel<ACT, NX, NT, false, true>)
                          : ((nw <= NC) ? (const void*)mlp_jet_bwd_kernel<ACT, NX, NT, true, false>
                                        : (const void*)mlp_jet_bwd_kernel<ACT, NX, NT, false, false>);
    return grid_for(k, nw * 32, smem, num_tiles, grid_out, spill);
}

template <int ACT, int NX, int NT>
int launch_bwd_impl(const BwdArgs& a, int grid, cudaStream_t stream) {
    constexpr int J = 1 + NX + NT;
    const int nw = bwd_block_warps(a.net.W, a.net.L, a.net.din, J);
    const bool spill = bwd_spills(a.net.W, a.net.L, a.net.din, J);
    if (spill != (a.hbuf != nullptr)) return PDEREAD_ERR_BAD_ARGUMENT;
    const size_t smem = spill ? bwd_smem_small(a.net.W, a.net.L, a.net.din, J) : bwd_smem_bytes(a.net.W, a.net.L, a.net.din, J);
    BwdArgs b = a;
    b.aux = (!spill && bwd_stage_weights(a.net.W, a.net.L, a.net.din, J)) ? 1 : 0;
    // (the dynamic shared-memory attribute was set by bwd_grid -> grid_for for this shape)
    if (spill) {
        if (nw <= a.net.NC) mlp_jet_bwd_kernel<ACT, NX, NT, true, true><<<grid, nw * 32, smem, stream>>>(b);
        else                mlp_jet_bwd_kernel<ACT, NX, NT, false, true><<<grid, nw * 32, smem, stream>>>(b);
    } else {
        if (nw <= a.net.NC) mlp_jet_bwd_kernel<ACT, NX, NT, true, false><<<grid, nw * 32, smem, stream>>>(b);
        else                mlp_jet_bwd_kernel<ACT, NX, NT, false, false><<<grid, nw * 32, smem, stream>>>(b);
    }
    PDR_CUDA_CHECK(cudaGetLastError());
    return PDEREAD_OK;
}

}  // namespace pdr
