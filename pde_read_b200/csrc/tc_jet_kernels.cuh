// Tensor-core (tcgen05 / TMEM) MLP-jet kernels for sm_100a.
//
// Same math as mlp_jet_kernels.cuh (reference Network.py:174-202 + the nested autograd of
// PDE_Residual.py:92-145, replaced by Taylor-jet propagation), but the hidden-layer products
//     z_j[i][p] = sum_k W_l[i][k] * y_j[k][p]          (i, k neurons; p point; j jet index)
// run on the 5th-generation tensor cores as D[M = neuron i][N = (point, jet)] = A[i][k] * B[k][(p,j)]:
//   * A = the layer's weight matrix, K-major in shared memory (streamed per layer by cp.async.bulk),
//   * B = the activations of the layer below, written by the epilogue warps straight from registers
//         into shared memory in the K-major core-matrix layout,
//   * D = fp32 accumulators in TMEM; the result of layer l+1 overwrites the columns of layer l in place.
// fp32 fidelity comes from the 3xTF32 split (hi*hi + lo*hi + hi*lo), see umma.cuh.
//
// Thread mapping of the epilogue (8 warps): TMEM lane = neuron, so one thread owns one neuron and walks
// over the points of a chunk: tcgen05.ld its J jet coefficients, add the bias, apply the activation
// jet (jet_math.cuh) in registers, split to hi/lo and store the B operand of the next layer.  A tile
// is two chunks that ping-pong between the epilogue warps and the single MMA-issuing thread through
// mbarriers (full[c]: operand written -> MMA may run; dready[c]: tcgen05.commit -> accumulators ready).
// Width <= 64 uses M = 64 MMAs and packs two independent point sets ("subs") into the two 16-lane
// halves of every TMEM sub-partition, so that all 32 lanes of a warp carry a neuron.
#pragma once
#include <cuda_runtime.h>

#include "internal.h"
#include "jet_math.cuh"
#include "mlp_jet_kernels.cuh"
#include "umma.cuh"

namespace pdr {

constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = (TC_EPI_WARPS + 2) * 32;   // + MMA warp + weight-loader warp
constexpr int TC_FLUSH_TILES = 2;   // the TMEM weight-gradient accumulators are drained every this many tiles
constexpr uint32_t TC_B_CS = 144;                      // byte stride between 4-k groups of the B operand (bank spread)

// ---- small PTX helpers not in umma.cuh -------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :
                 : "r"(umma::smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(umma::smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
                 :
                 : "r"(umma::smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;\n" : : "n"(TC_EPI_WARPS * 32) : "memory"); }

// N consecutive TMEM columns of this thread's lane -> r[0..N) (no wait; call tmem_wait_ld() before use)
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* r) {
    if constexpr (N >= 16) {
        umma::tmem_ld_x16(taddr, r);
        tmem_ld_cols<N - 16>(taddr + 16, r + 16);
    } else if constexpr (N >= 8) {
        umma::tmem_ld_x8(taddr, r);
        tmem_ld_cols<N - 8>(taddr + 8, r + 8);
    } else if constexpr (N >= 4) {
        umma::tmem_ld_x4(taddr, r);
        tmem_ld_cols<N - 4>(taddr + 4, r + 4);
    } else if constexpr (N >= 2) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
        tmem_ld_cols<N - 2>(taddr + 2, r + 2);
    } else if constexpr (N == 1) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=r"(r[0]) : "r"(taddr) : "memory");
    }
}

// ---- forward ------------------------------------------------------------------------------------
template <int ACT, int NX, int NT, int MROWS, bool STASH>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_jet_fwd_kernel(const TcFwdArgs a) {
    constexpr int J = 1 + NX + NT;
    constexpr int SUBS = (MROWS == 64) ? 2 : 1;
    constexpr int G = (J <= 4) ? 4 : 2;       // points per TMEM load / independent activation chains per thread
    using namespace umma;

    extern __shared__ __align__(128) unsigned char tc_smem[];
    unsigned char* smem = tc_smem;
    const NetDev& net = a.net;
    const TcGeom& g = a.geom;
    const int W = net.W, L = net.L, din = net.din;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TP = g.TP, PC = g.PC;

    unsigned char* wbuf = smem;                                               // [nwbuf][hi|lo][a_bytes]
    unsigned char* slots = wbuf + (size_t)g.nwbuf * 2 * g.a_bytes;            // [SUBS][2 chunks][hi|lo][slot_bytes]
    float* IN = reinterpret_cast<float*>(slots + (size_t)SUBS * g.sub_stride);   // [din][TP]
    uint64_t* bars = reinterpret_cast<uint64_t*>(IN + ((din * TP + 1) & ~1));
    uint64_t* full = bars;          // [2]  epilogue -> MMA   (count TC_EPI_WARPS)
    uint64_t* dready = bars + 2;    // [2]  MMA -> epilogue   (tcgen05.commit)
    uint64_t* wfull = bars + 4;     // [2]  weights landed    (expect_tx)
    uint64_t* wfree = bars + 6;     // [2]  weights consumed  (tcgen05.commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    // one-time setup: zero the operand slots (their k / n padding must stay finite), barriers, TMEM
    {
        uint4* p = reinterpret_cast<uint4*>(slots);
        const int n16 = (SUBS * g.sub_stride) >> 4;
        for (int i = threadIdx.x; i < n16; i += blockDim.x) p[i] = make_uint4(0, 0, 0, 0);
    }
    if (threadIdx.x == 0) {
        for (int c = 0; c < 2; ++c) {
            mbar_init(&full[c], TC_EPI_WARPS);
            mbar_init(&dready[c], 1);
            mbar_init(&wfull[c], 1);
            mbar_init(&wfree[c], 1);
        }
        fence_mbar_init();
    }
    if (warp == TC_EPI_WARPS) tmem_alloc(tmem_slot, (uint32_t)g.tmem_cols);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const long long my_tiles = (a.num_tiles > blockIdx.x) ? (a.num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long total_gemms = my_tiles * L;     // L-1 hidden products + the output product per tile
    const uint32_t img_bytes = 2u * (uint32_t)g.a_bytes;

    if (warp == TC_EPI_WARPS + 1) {
        // ===== weight loader: one bulk copy (hi|lo image of one layer) per product =====
        if (lane == 0) {
            // L <= nwbuf: every product keeps its own buffer, the images are loaded once and stay resident
            const long long nload = (L <= g.nwbuf) ? (total_gemms < L ? total_gemms : L) : total_gemms;
            for (long long s = 0; s < nload; ++s) {
                const int b = (int)(s % g.nwbuf);
                if (s >= g.nwbuf) mbar_wait(&wfree[b], (uint32_t)(((s / g.nwbuf) - 1) & 1));
                const int gi = (int)(s % L);        // product index inside the tile: 0..L-2 hidden, L-1 output
                mbar_expect_tx(&wfull[b], img_bytes);
                bulk_g2s(wbuf + (size_t)b * img_bytes, reinterpret_cast<const unsigned char*>(a.wimg) + (size_t)gi * img_bytes,
                         img_bytes, &wfull[b]);
            }
        }
    } else if (warp == TC_EPI_WARPS) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(MROWS, g.NCpad, 0, 0);
            const uint32_t corr_off = 2u * (uint32_t)g.NCpad;
            const int ksteps = g.Kpad >> 3;
            uint32_t full_ph[2] = {0, 0};
            const bool resident = (L <= g.nwbuf);
            for (long long s = 0; s < total_gemms; ++s) {
                const int b = resident ? (int)(s % L) : (int)(s % g.nwbuf);
                if (!resident || s < L) mbar_wait(&wfull[b], (uint32_t)((s / g.nwbuf) & 1));
                const uint32_t a_hi = smem_u32(wbuf + (size_t)b * img_bytes), a_lo = a_hi + (uint32_t)g.a_bytes;
                for (int c = 0; c < 2; ++c) {
                    mbar_wait(&full[c], full_ph[c]);
                    full_ph[c] ^= 1;
                    tc_fence_after();
#pragma unroll
                    for (int sub = 0; sub < SUBS; ++sub) {
                        const uint32_t d_addr = tmem_base + ((uint32_t)(sub * 16) << 16) + (uint32_t)(c * g.NCpad);
                        const uint32_t b_hi = smem_u32(slots + (size_t)sub * g.sub_stride + (size_t)(c * 2) * g.slot_bytes);
                        const uint32_t b_lo = b_hi + (uint32_t)g.slot_bytes;
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint64_t dah = make_smem_desc(a_hi + ks * 256u, 128u, (uint32_t)g.a_rs);
                            const uint64_t dal = make_smem_desc(a_lo + ks * 256u, 128u, (uint32_t)g.a_rs);
                            const uint64_t dbh = make_smem_desc(b_hi + ks * 2u * TC_B_CS, TC_B_CS, (uint32_t)g.b_rs);
                            const uint64_t dbl = make_smem_desc(b_lo + ks * 2u * TC_B_CS, TC_B_CS, (uint32_t)g.b_rs);
                            // the tensor core truncates when it accumulates: keep the large hi*hi sum and the small
                            // correction terms in separate accumulators; the epilogue adds them in fp32 (round to nearest)
                            mma_tf32_ss(d_addr, dah, dbh, idesc, ks > 0);
                            mma_tf32_ss(d_addr + corr_off, dal, dbh, idesc, ks > 0);
                            mma_tf32_ss(d_addr + corr_off, dah, dbl, idesc, 1);
                        }
                    }
                    mma_commit(&dready[c]);
                }
                if (!resident) mma_commit(&wfree[b]);
            }
        }
    } else {
        // ===== epilogue warps =====
        const int qd = warp & 3, half = warp >> 2;
        const int sub = (MROWS == 64) ? (lane >> 4) : 0;
        const int rl = (MROWS == 64) ? (lane & 15) : lane;                 // row inside the warp's lane block
        const int row = (MROWS == 64) ? (16 * qd + rl) : (32 * qd + rl);   // row of the weight image / D
        const int nid = qd * g.NB + rl;                                    // neuron owned by this thread
        const bool active = (rl < g.NB) && (nid < W);
        const bool out_thread = (qd == 0) && (rl == 0);                    // row 0 of the output product
        const uint32_t lane_taddr = tmem_base + ((uint32_t)(32 * qd) << 16);
        const uint32_t koff = (uint32_t)(nid & 3) * 4u + (uint32_t)(nid >> 2) * TC_B_CS;
        const int PH = PC >> 1;                                            // points per thread per chunk
        const int et = threadIdx.x;                                        // 0 .. 255
        uint32_t dr_ph[2] = {0, 0};
        double lsum = 0.0;
        const float bout = __ldg(net.b[L]);

        for (long long tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
            const long long p0 = tile * TP;
            for (int idx = et; idx < din * TP; idx += TC_EPI_WARPS * 32) {
                const int pt = idx / din, d = idx - pt * din;
                const long long p = p0 + pt;
                IN[d * TP + pt] = (p < a.M) ? __ldg(a.in + p * din + d) : 0.f;
            }
            epi_bar_sync();

            // ---- hidden layer 0 on the CUDA cores (K = din is tiny) ----
            {
                float ab[7];
                load_ab<ACT>(net, 0, ab);
                for (int c = 0; c < 2; ++c) {
                    unsigned char* s_hi = slots + (size_t)sub * g.sub_stride + (size_t)(c * 2) * g.slot_bytes;
                    unsigned char* s_lo = s_hi + g.slot_bytes;
                    if (active) {
                        for (int q = 0; q < PH; ++q) {
                            const int pp = half * PH + q;                      // point inside the chunk
                            const int pt = (sub * 2 + c) * PC + pp;            // point inside the tile
                            float z[J], y[J];
                            input_layer_jet<NX, NT>(net, IN, TP, nid, pt, z);
                            act_fwd<ACT, NX, NT, float>(z, ab, y);
#pragma unroll
                            for (int j = 0; j < J; ++j) {
                                const int n = pp * J + j;
                                const uint32_t off = (uint32_t)(n & 7) * 16u + (uint32_t)(n >> 3) * (uint32_t)g.b_rs + koff;
                                float hi, lo;
                                tf32_split(y[j], hi, lo);
                                *reinterpret_cast<float*>(s_hi + off) = hi;
                                *reinterpret_cast<float*>(s_lo + off) = lo;
                            }
                        }
                    }
                    fence_proxy_async();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[c]);
                }
            }

            // ---- products 1 .. L: hidden layers 1..L-1, then the output layer ----
            for (int gi = 1; gi <= L; ++gi) {
                const bool is_out = (gi == L);
                float ab[7];
                float bias = 0.f;
                if (!is_out) {
                    load_ab<ACT>(net, gi, ab);
                    if (active) bias = __ldg(net.b[gi] + nid);
                }
                for (int c = 0; c < 2; ++c) {
                    mbar_wait(&dready[c], dr_ph[c]);
                    dr_ph[c] ^= 1;
                    tc_fence_after();
                    unsigned char* s_hi = slots + (size_t)sub * g.sub_stride + (size_t)(c * 2) * g.slot_bytes;
                    unsigned char* s_lo = s_hi + g.slot_bytes;
                    const uint32_t col0 = (uint32_t)(c * g.NCpad);
                    for (int q0 = 0; q0 < PH; q0 += G) {
                        uint32_t raw[G * J], rawc[G * J];
                        tmem_ld_cols<G * J>(lane_taddr + col0 + (uint32_t)((half * PH + q0) * J), raw);
                        tmem_ld_cols<G * J>(lane_taddr + col0 + 2u * (uint32_t)g.NCpad + (uint32_t)((half * PH + q0) * J), rawc);
                        tmem_wait_ld();
#pragma unroll
                        for (int gq = 0; gq < G; ++gq) {
                            const int pp = half * PH + q0 + gq;
                            const int pt = (sub * 2 + c) * PC + pp;
                            float z[J];
#pragma unroll
                            for (int j = 0; j < J; ++j) z[j] = __uint_as_float(raw[gq * J + j]) + __uint_as_float(rawc[gq * J + j]);
                            if (!is_out) {
                                z[0] += bias;
                                if (STASH && p0 + pt < a.M) {
                                    // [layer][point][jet][row]: consecutive lanes -> consecutive addresses
                                    float* sp = a.stash + ((size_t)(gi - 1) * a.M + (size_t)(p0 + pt)) * (size_t)(J * MROWS) + row;
#pragma unroll
                                    for (int j = 0; j < J; ++j) sp[j * MROWS] = z[j];
                                }
                                if (active) {
                                    float y[J];
                                    act_fwd<ACT, NX, NT, float>(z, ab, y);
#pragma unroll
                                    for (int j = 0; j < J; ++j) {
                                        const int n = pp * J + j;
                                        const uint32_t off = (uint32_t)(n & 7) * 16u + (uint32_t)(n >> 3) * (uint32_t)g.b_rs + koff;
                                        float hi, lo;
                                        tf32_split(y[j], hi, lo);
                                        *reinterpret_cast<float*>(s_hi + off) = hi;
                                        *reinterpret_cast<float*>(s_lo + off) = lo;
                                    }
                                }
                            } else if (out_thread) {
                                const long long p = p0 + pt;
                                if (p < a.M) {
                                    const float u0 = z[0] + bout;
                                    if (J == 1) {
                                        if (a.dtm_in != nullptr) {
                                            const float R = __ldg(a.dtm_in + p) - u0;
                                            if (a.resid_out) a.resid_out[p] = R;
                                            if (a.seed_out)
                                                a.seed_out[p] = (a.grad_resid != nullptr) ? -__ldg(a.grad_resid + p) : -a.seed_scale * R;
                                            lsum += (double)R * (double)R;
                                        } else {
                                            a.out0[p] = u0;
                                        }
                                    } else {
                                        float f = 1.f;
#pragma unroll
                                        for (int j = 0; j <= NX; ++j) {
                                            if (j >= 2) f *= (float)j;
                                            a.out0[p * (NX + 1) + j] = (j == 0 ? u0 : z[j]) * f;
                                        }
                                        if (NT > 0 && a.out1 != nullptr) {
                                            float ft = 1.f;
#pragma unroll
                                            for (int t = 2; t <= NT; ++t) ft *= (float)t;
                                            a.out1[p] = z[NX + NT] * ft;
                                        }
                                    }
                                }
                            }
                        }
                    }
                    if (!is_out) {
                        fence_proxy_async();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&full[c]);
                    } else {
                        tc_fence_before();   // orders these TMEM reads before the next tile's arrive on full[c]
                    }
                }
            }
        }

        if (J == 1 && a.sumsq != nullptr && out_thread && lsum != 0.0) atomicAdd(a.sumsq, lsum);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == TC_EPI_WARPS) tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
}


// ---- backward -------------------------------------------------------------------------------------
// Reverse pass of tc_jet_fwd_kernel for one network.  Per hidden layer lv (top first) the epilogue
// warps turn ybar^lv (from the output seeds, or from TMEM) into zbar^lv with the hand-written
// activation VJP and re-evaluate y^(lv-1) from the stashed pre-activations; both go to shared memory
// as tensor-core operands and the MMA thread issues
//     hbar^(lv-1)[k][n] = sum_i W_lv[i][k] zbar^lv[i][n]          (A = W_lv^T image, B = zbar [n][i])
//     gW_lv[i][k]      += sum_n zbar^lv[i][n] y^(lv-1)[k][n]      (A = zbar [i][n],   B = y [k][n])
// The weight-gradient accumulators stay in TMEM for the whole kernel and are flushed once at the end;
// bias / input-layer / output-layer / rational-coefficient gradients are thread-local sums (thread =
// neuron) collected in shared memory.
template <int J>
struct TcCols {
    static constexpr int Jp = (J == 3) ? 4 : (J == 5) ? 6 : (J == 7) ? 8 : J;   // columns per point
    static constexpr int GB = (Jp == 1) ? 8 : (Jp == 2) ? 4 : 2;                 // points per group (backward)
};

template <int ACT, int NX, int NT, int MROWS>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_jet_bwd_kernel(const TcBwdArgs a) {
    constexpr int J = 1 + NX + NT;
    constexpr int Jp = TcCols<J>::Jp;
    constexpr int G = TcCols<J>::GB;
    constexpr int GC = G * Jp;                 // columns per group (multiple of 4)
    constexpr int SUBS = (MROWS == 64) ? 2 : 1;
    using namespace umma;

    extern __shared__ __align__(128) unsigned char tc_smem_b[];
    unsigned char* smem = tc_smem_b;
    const NetDev& net = a.net;
    const TcBwdGeom& g = a.geom;
    const int W = net.W, L = net.L, din = net.din;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TP = g.TP, PC = g.PC, NCH = g.nchunk;
    const bool want_gin = (a.gin != nullptr);

    const int sbL = L * W, sw0 = sbL + 1, swout = sw0 + W * din, sab = swout + W, NS = sab + 7 * L;
    unsigned char* wbuf = smem;
    unsigned char* slots = wbuf + (size_t)g.nwbuf * 2 * g.a_bytes;
    float* IN = reinterpret_cast<float*>(slots + (size_t)SUBS * g.sub_stride);   // [din][TP]
    float* SD = IN + din * TP;                                                    // [J][TP]
    float* SACC = SD + J * TP;                                                    // [NS]
    uint64_t* bars = reinterpret_cast<uint64_t*>(SACC + ((NS + 1) & ~1));
    uint64_t* full = bars;
    uint64_t* dready = bars + 2;
    uint64_t* wfull = bars + 4;
    uint64_t* wfree = bars + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    float* gp = a.gpart + (size_t)blockIdx.x * a.lay.total;
    {
        uint4* p = reinterpret_cast<uint4*>(slots);
        const int n16 = (SUBS * g.sub_stride) >> 4;
        for (int i = threadIdx.x; i < n16; i += blockDim.x) p[i] = make_uint4(0, 0, 0, 0);
        for (int i = threadIdx.x; i < NS; i += blockDim.x) SACC[i] = 0.f;
        if (!a.accumulate)
            for (int i = threadIdx.x; i < a.lay.total; i += blockDim.x) gp[i] = 0.f;
    }
    if (threadIdx.x == 0) {
        for (int c = 0; c < 2; ++c) {
            mbar_init(&full[c], TC_EPI_WARPS);
            mbar_init(&dready[c], 1);
            mbar_init(&wfull[c], 1);
            mbar_init(&wfree[c], 1);
        }
        fence_mbar_init();
    }
    if (warp == TC_EPI_WARPS) tmem_alloc(tmem_slot, (uint32_t)g.tmem_cols);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const long long my_tiles = (a.num_tiles > blockIdx.x) ? (a.num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int nprod = (L - 1) + (want_gin ? 1 : 0);
    const long long total_prod = my_tiles * nprod;
    const uint32_t img_bytes = 2u * (uint32_t)g.a_bytes;
    const uint32_t half_set = (uint32_t)(g.zt_bytes + g.za_bytes + g.yb_bytes);   // hi part of a set; lo follows

    if (warp == TC_EPI_WARPS + 1) {
        // ===== weight loader =====
        if (lane == 0) {
            const long long nload = (nprod <= g.nwbuf) ? (total_prod < nprod ? total_prod : nprod) : total_prod;
            for (long long s = 0; s < nload; ++s) {
                const int b = (int)(s % g.nwbuf);
                if (s >= g.nwbuf) mbar_wait(&wfree[b], (uint32_t)(((s / g.nwbuf) - 1) & 1));
                const int idx = (int)(s % nprod);
                const int img = (idx < L - 1) ? (L - 2 - idx) : (L - 1);
                mbar_expect_tx(&wfull[b], img_bytes);
                bulk_g2s(wbuf + (size_t)b * img_bytes, reinterpret_cast<const unsigned char*>(a.wimg) + (size_t)img * img_bytes,
                         img_bytes, &wfull[b]);
            }
        }
    } else if (warp == TC_EPI_WARPS) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc_d = make_idesc_tf32(MROWS, g.NCpad, 0, 0);
            const uint32_t idesc_w = make_idesc_tf32(MROWS, g.npad_w, 0, 0);
            const int ks_d = g.Kpad >> 3, ks_w = g.NC >> 3;
            const uint32_t corr_off = (uint32_t)(NCH * g.NCpad);   // correction-term accumulators of the hbar products
            uint32_t full_ph[2] = {0, 0};
            for (long long s = 0; s < total_prod; ++s) {
                const int idx = (int)(s % nprod);
                const int b = (nprod <= g.nwbuf) ? idx : (int)(s % g.nwbuf);
                const bool first_tile = ((s / nprod) % TC_FLUSH_TILES) == 0;   // accumulators were flushed: restart them
                const bool resident = (nprod <= g.nwbuf);
                if (!resident || s < nprod) mbar_wait(&wfull[b], (uint32_t)((s / g.nwbuf) & 1));
                const uint32_t a_hi = smem_u32(wbuf + (size_t)b * img_bytes), a_lo = a_hi + (uint32_t)g.a_bytes;
                for (int c = 0; c < NCH; ++c) {
                    mbar_wait(&full[c], full_ph[c]);
                    full_ph[c] ^= 1;
                    tc_fence_after();
#pragma unroll
                    for (int sub = 0; sub < SUBS; ++sub) {
                        const uint32_t set_hi = smem_u32(slots + (size_t)sub * g.sub_stride + (size_t)c * g.set_bytes);
                        const uint32_t set_lo = set_hi + half_set;
                        const uint32_t lane_off = (uint32_t)(sub * 16) << 16;
                        // hbar (or the input gradient): D[k][n] = sum_i A[k][i] * zbar[n][i]
                        {
                            const uint32_t d_addr = tmem_base + lane_off + (uint32_t)(c * g.NCpad);
                            for (int ks = 0; ks < ks_d; ++ks) {
                                const uint64_t dah = make_smem_desc(a_hi + ks * 256u, 128u, (uint32_t)g.a_rs);
                                const uint64_t dal = make_smem_desc(a_lo + ks * 256u, 128u, (uint32_t)g.a_rs);
                                const uint64_t dbh = make_smem_desc(set_hi + ks * 2u * TC_B_CS, TC_B_CS, (uint32_t)g.zt_rs);
                                const uint64_t dbl = make_smem_desc(set_lo + ks * 2u * TC_B_CS, TC_B_CS, (uint32_t)g.zt_rs);
                                mma_tf32_ss(d_addr, dah, dbh, idesc_d, ks > 0);
                                mma_tf32_ss(d_addr + corr_off, dal, dbh, idesc_d, ks > 0);
                                mma_tf32_ss(d_addr + corr_off, dah, dbl, idesc_d, 1);
                            }
                        }
                        if (idx < L - 1) {
                            // gW_lv[i][k] += sum_n zbar[i][n] * y[k][n]
                            const int lv = L - 1 - idx;
                            const uint32_t d_addr = tmem_base + lane_off + (uint32_t)(g.gcol0 + (lv - 1) * g.npad_w);
                            const uint32_t za_hi = set_hi + (uint32_t)g.zt_bytes, za_lo = set_lo + (uint32_t)g.zt_bytes;
                            const uint32_t yb_hi = za_hi + (uint32_t)g.za_bytes, yb_lo = za_lo + (uint32_t)g.za_bytes;
                            for (int ks = 0; ks < ks_w; ++ks) {
                                const uint64_t dah = make_smem_desc(za_hi + ks * 256u, 128u, (uint32_t)g.za_rs);
                                const uint64_t dal = make_smem_desc(za_lo + ks * 256u, 128u, (uint32_t)g.za_rs);
                                const uint64_t dbh = make_smem_desc(yb_hi + ks * 256u, 128u, (uint32_t)g.za_rs);
                                const uint64_t dbl = make_smem_desc(yb_lo + ks * 256u, 128u, (uint32_t)g.za_rs);
                                mma_tf32_ss(d_addr, dah, dbh, idesc_w, !(first_tile && c == 0 && ks == 0));
                                mma_tf32_ss(d_addr, dal, dbh, idesc_w, 1);
                                mma_tf32_ss(d_addr, dah, dbl, idesc_w, 1);
                            }
                        }
                    }
                    mma_commit(&dready[c]);
                }
                if (!resident) mma_commit(&wfree[b]);
            }
        }
    } else {
        // ===== epilogue warps =====
        const int qd = warp & 3, half = warp >> 2;
        const int sub = (MROWS == 64) ? (lane >> 4) : 0;
        const int rl = (MROWS == 64) ? (lane & 15) : lane;
        const int row = (MROWS == 64) ? (16 * qd + rl) : (32 * qd + rl);
        const int nid = qd * g.NB + rl;
        const bool active = (rl < g.NB) && (nid < W);
        const uint32_t lane_taddr = tmem_base + ((uint32_t)(32 * qd) << 16);
        const uint32_t koff = (uint32_t)(nid & 3) * 4u + (uint32_t)(nid >> 2) * TC_B_CS;
        const uint32_t za_row = (uint32_t)(row & 7) * 16u + (uint32_t)(row >> 3) * (uint32_t)g.za_rs;
        const uint32_t yb_row = (uint32_t)(nid & 7) * 16u + (uint32_t)(nid >> 3) * (uint32_t)g.za_rs;
        const int PH = PC >> 1;
        const int et = threadIdx.x;
        uint32_t dr_ph[2] = {0, 0};
        const float* wout = net.w[L];
        const size_t srow = (size_t)(J * MROWS);
        long long tiles_done = 0;

        for (long long tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
            const long long p0 = tile * TP;
            for (int idx = et; idx < din * TP; idx += TC_EPI_WARPS * 32) {
                const int pt = idx / din, d = idx - pt * din;
                const long long p = p0 + pt;
                IN[d * TP + pt] = (p < a.M) ? __ldg(a.in + p * din + d) : 0.f;
            }
            for (int idx = et; idx < J * TP; idx += TC_EPI_WARPS * 32) {
                const int j = idx / TP, pt = idx - j * TP;
                const long long p = p0 + pt;
                float sd = 0.f;
                if (p < a.M) {
                    if (J == 1) {
                        sd = a.g0 ? a.g0_scale * __ldg(a.g0 + p) : 0.f;
                    } else if (j <= NX) {
                        float f = 1.f;
#pragma unroll
                        for (int t = 2; t <= NX; ++t) if (t <= j) f *= (float)t;
                        sd = a.g0 ? a.g0_scale * f * __ldg(a.g0 + p * (NX + 1) + j) : 0.f;
                    } else if (j == NX + NT) {
                        float f = 1.f;
#pragma unroll
                        for (int t = 2; t <= NT; ++t) f *= (float)t;
                        sd = a.g1 ? a.g1_scale * f * __ldg(a.g1 + p) : 0.f;
                    }
                }
                SD[idx] = sd;
            }
            epi_bar_sync();
            if (warp == 0) {   // output bias: sum of the order-0 seeds
                float sb = 0.f;
                for (int pt = lane; pt < TP; pt += 32) sb += SD[pt];
                sb = warp_sum(sb);
                if (lane == 0) SACC[sbL] += sb;
            }

            for (int lv = L - 1; lv >= 0; --lv) {
                const bool top = (lv == L - 1);
                const bool feeds_mma = (lv >= 1) || want_gin;
                float ab[7], ab_lo[7], abbar[7];
                load_ab<ACT>(net, lv, ab);
                load_ab<ACT>(net, lv >= 1 ? lv - 1 : 0, ab_lo);
#pragma unroll
                for (int t = 0; t < 7; ++t) abbar[t] = 0.f;
                const float wo = (top && active) ? __ldg(wout + nid) : 0.f;
                const float* st_z = (lv >= 1) ? a.stash + ((size_t)(lv - 1) * a.M + (size_t)p0) * srow + row : nullptr;
                const float* st_zp = (lv >= 2) ? a.stash + ((size_t)(lv - 2) * a.M + (size_t)p0) * srow + row : nullptr;

                for (int c = 0; c < NCH; ++c) {
                    // stash reads do not depend on the tensor core: issue the first group's before waiting for it
                    float zn[G][J], zpn[G][J];
                    auto fetch = [&](int q0f) {
#pragma unroll
                        for (int gq = 0; gq < G; ++gq) {
                            const int pt = (sub * NCH + c) * PC + half * PH + q0f + gq;
                            const bool valid = active && (p0 + pt) < a.M;
#pragma unroll
                            for (int j = 0; j < J; ++j) {
                                zn[gq][j] = (lv >= 1 && valid) ? __ldcs(st_z + (size_t)pt * srow + j * MROWS) : 0.f;
                                zpn[gq][j] = (lv >= 2 && valid) ? __ldcs(st_zp + (size_t)pt * srow + j * MROWS) : 0.f;
                            }
                        }
                    };
                    fetch(0);
                    if (!top) {
                        mbar_wait(&dready[c], dr_ph[c]);
                        dr_ph[c] ^= 1;
                        tc_fence_after();
                    }
                    unsigned char* set_hi = slots + (size_t)sub * g.sub_stride + (size_t)c * g.set_bytes;
                    unsigned char* set_lo = set_hi + half_set;
                    float gb = 0.f, gwo = 0.f, gx = 0.f, gt = 0.f;
                    float gwin[MAXD];
#pragma unroll
                    for (int d = 0; d < MAXD; ++d) gwin[d] = 0.f;

                    for (int q0 = 0; q0 < PH; q0 += G) {
                        const int pp0 = half * PH + q0;
                        float zc[G][J], zpc[G][J];
#pragma unroll
                        for (int gq = 0; gq < G; ++gq)
#pragma unroll
                            for (int j = 0; j < J; ++j) { zc[gq][j] = zn[gq][j]; zpc[gq][j] = zpn[gq][j]; }
                        if (q0 + G < PH) fetch(q0 + G);
                        uint32_t raw[GC], rawc[GC];
                        if (!top) {
                            tmem_ld_cols<GC>(lane_taddr + (uint32_t)(c * g.NCpad + pp0 * Jp), raw);
                            tmem_ld_cols<GC>(lane_taddr + (uint32_t)((NCH + c) * g.NCpad + pp0 * Jp), rawc);
                            tmem_wait_ld();
                        }
                        float zb[GC], yp[GC];
#pragma unroll
                        for (int t = 0; t < GC; ++t) { zb[t] = 0.f; yp[t] = 0.f; }
                        if (active) {
#pragma unroll
                            for (int gq = 0; gq < G; ++gq) {
                                const int pt = (sub * NCH + c) * PC + pp0 + gq;
                                float z[J], ybar[J], zbar[J];
                                if (lv == 0) {
                                    input_layer_jet<NX, NT>(net, IN, TP, nid, pt, z);
                                } else {
#pragma unroll
                                    for (int j = 0; j < J; ++j) z[j] = zc[gq][j];
                                }
                                if (top) {
                                    float y[J];
                                    act_fwd<ACT, NX, NT, float>(z, ab, y);
#pragma unroll
                                    for (int j = 0; j < J; ++j) {
                                        const float sd = SD[j * TP + pt];
                                        ybar[j] = wo * sd;
                                        gwo = fmaf(sd, y[j], gwo);
                                    }
                                } else {
#pragma unroll
                                    for (int j = 0; j < J; ++j)
                                        ybar[j] = __uint_as_float(raw[gq * Jp + j]) + __uint_as_float(rawc[gq * Jp + j]);
                                }
                                act_vjp<ACT, NX, NT, float>(z, ab, ybar, zbar, abbar);
                                gb += zbar[0];
#pragma unroll
                                for (int j = 0; j < J; ++j) zb[gq * Jp + j] = zbar[j];
                                if (lv == 0) {
                                    if (NX > 0) gx += zbar[1];
                                    if (NT > 0) gt += zbar[NX + 1];
#pragma unroll
                                    for (int d = 0; d < MAXD; ++d)
                                        if (d < din) gwin[d] = fmaf(zbar[0], IN[d * TP + pt], gwin[d]);
                                } else {
                                    float zp[J], y1[J];
                                    if (lv == 1) {
                                        input_layer_jet<NX, NT>(net, IN, TP, nid, pt, zp);
                                    } else {
#pragma unroll
                                        for (int j = 0; j < J; ++j) zp[j] = zpc[gq][j];
                                    }
                                    act_fwd<ACT, NX, NT, float>(zp, ab_lo, y1);
#pragma unroll
                                    for (int j = 0; j < J; ++j) yp[gq * Jp + j] = y1[j];
                                }
                            }
                            if (feeds_mma) {
                                // operand stores: zbar as [n][k] (scalar) and, for the weight gradient, zbar [row][n] and
                                // y [k][n] (16-byte vectors: the group's GC columns are contiguous and 4-aligned)
                                const int n0 = pp0 * Jp;
                                float zh[GC], zl[GC];
#pragma unroll
                                for (int t = 0; t < GC; ++t) tf32_split(zb[t], zh[t], zl[t]);
#pragma unroll
                                for (int t = 0; t < GC; ++t) {
                                    if ((t % Jp) < J) {
                                        const int n = n0 + t;
                                        const uint32_t off = (uint32_t)(n & 7) * 16u + (uint32_t)(n >> 3) * (uint32_t)g.zt_rs + koff;
                                        *reinterpret_cast<float*>(set_hi + off) = zh[t];
                                        *reinterpret_cast<float*>(set_lo + off) = zl[t];
                                    }
                                }
                                if (lv >= 1) {
                                    unsigned char* za_h = set_hi + g.zt_bytes + za_row + (uint32_t)(n0 >> 2) * 128u;
                                    unsigned char* za_l = set_lo + g.zt_bytes + za_row + (uint32_t)(n0 >> 2) * 128u;
                                    unsigned char* yb_h = set_hi + g.zt_bytes + g.za_bytes + yb_row + (uint32_t)(n0 >> 2) * 128u;
                                    unsigned char* yb_l = set_lo + g.zt_bytes + g.za_bytes + yb_row + (uint32_t)(n0 >> 2) * 128u;
#pragma unroll
                                    for (int t4 = 0; t4 < GC / 4; ++t4) {
                                        *reinterpret_cast<float4*>(za_h + t4 * 128) = make_float4(zh[4 * t4], zh[4 * t4 + 1], zh[4 * t4 + 2], zh[4 * t4 + 3]);
                                        *reinterpret_cast<float4*>(za_l + t4 * 128) = make_float4(zl[4 * t4], zl[4 * t4 + 1], zl[4 * t4 + 2], zl[4 * t4 + 3]);
                                        float yh[4], yl[4];
#pragma unroll
                                        for (int u = 0; u < 4; ++u) tf32_split(yp[4 * t4 + u], yh[u], yl[u]);
                                        *reinterpret_cast<float4*>(yb_h + t4 * 128) = make_float4(yh[0], yh[1], yh[2], yh[3]);
                                        *reinterpret_cast<float4*>(yb_l + t4 * 128) = make_float4(yl[0], yl[1], yl[2], yl[3]);
                                    }
                                }
                            }
                        }
                    }
                    if (active) {
                        atomicAdd(&SACC[lv * W + nid], gb);
                        if (top) atomicAdd(&SACC[swout + nid], gwo);
                        if (lv == 0) {
                            if (NX > 0) gwin[1] += gx;
                            if (NT > 0) gwin[0] += gt;
#pragma unroll
                            for (int d = 0; d < MAXD; ++d)
                                if (d < din) atomicAdd(&SACC[sw0 + nid * din + d], gwin[d]);
                        }
                    }
                    if (feeds_mma) {
                        fence_proxy_async();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&full[c]);
                    } else {
                        tc_fence_before();
                    }
                }
                if (ACT == ACT_RATIONAL) {
#pragma unroll
                    for (int t = 0; t < 7; ++t) {
                        const float sv = warp_sum(abbar[t]);
                        if (lane == 0) atomicAdd(&SACC[sab + 7 * lv + t], sv);
                    }
                }
            }

            if (want_gin) {
                // input gradient (plain networks only): rows d < din of the last product
                for (int c = 0; c < NCH; ++c) {
                    mbar_wait(&dready[c], dr_ph[c]);
                    dr_ph[c] ^= 1;
                    tc_fence_after();
                    if (qd == 0) {
                        for (int q = 0; q < PH; ++q) {
                            const int pp = half * PH + q;
                            uint32_t r1[1], r2[1];
                            tmem_ld_cols<1>(lane_taddr + (uint32_t)(c * g.NCpad + pp * Jp), r1);
                            tmem_ld_cols<1>(lane_taddr + (uint32_t)((NCH + c) * g.NCpad + pp * Jp), r2);
                            tmem_wait_ld();
                            const int pt = (sub * NCH + c) * PC + pp;
                            const long long p = p0 + pt;
                            if (rl < din && p < a.M) a.gin[p * din + rl] = __uint_as_float(r1[0]) + __uint_as_float(r2[0]);
                        }
                    }
                    tc_fence_before();
                }
            }

            // ---- drain the TMEM weight-gradient accumulators every TC_FLUSH_TILES tiles (and after the last
            //      tile): long fp32 accumulation chains inside the tensor core drift, these adds round to nearest.
            //      All products of this tile have completed (every warp waited for their commits above). ----
            ++tiles_done;
            if ((tiles_done % TC_FLUSH_TILES) == 0 || tile + gridDim.x >= a.num_tiles) {
                tc_fence_after();
                if (half == 0) {
                    for (int l = 1; l < L; ++l) {
                        const uint32_t gcol = (uint32_t)(g.gcol0 + (l - 1) * g.npad_w);
                        for (int k0 = 0; k0 < g.Kpad; k0 += 8) {
                            uint32_t v[8];
                            tmem_ld_x8(lane_taddr + gcol + (uint32_t)k0, v);
                            tmem_wait_ld();
#pragma unroll
                            for (int t = 0; t < 8; ++t) {
                                float f = __uint_as_float(v[t]);
                                if (MROWS == 64) f += __shfl_xor_sync(0xffffffffu, f, 16);
                                const int k = k0 + t;
                                if (active && sub == 0 && k < W) atomicAdd(&gp[a.lay.off_w[l] + nid * W + k], f);   // RED: no round trip
                            }
                        }
                    }
                }
                tc_fence_before();
            }
        }

        // ---- flush the small-parameter sums (shared memory) ----
        epi_bar_sync();
        for (int idx = et; idx < NS; idx += TC_EPI_WARPS * 32) {
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

    tc_fence_before();
    __syncthreads();
    if (warp == TC_EPI_WARPS) tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
}

// ---- launcher -----------------------------------------------------------------------------------
template <int ACT, int NX, int NT>
int launch_tc_fwd(const TcFwdArgs& a, cudaStream_t stream) {
    int dev = 0, sms = 0;
    PDR_CUDA_CHECK(cudaGetDevice(&dev));
    PDR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long grid = sms;
    if (grid > a.num_tiles) grid = a.num_tiles;
    if (grid < 1) grid = 1;
    const size_t smem = a.geom.smem_bytes;
#define PDR_TC_LAUNCH(MR, ST)                                                                                            \
    do {                                                                                                                 \
        const int rc_ = ensure_dynamic_smem((const void*)tc_jet_fwd_kernel<ACT, NX, NT, MR, ST>, smem);                  \
        if (rc_ != PDEREAD_OK) return rc_;                                                                               \
        tc_jet_fwd_kernel<ACT, NX, NT, MR, ST><<<(int)grid, TC_THREADS, smem, stream>>>(a);                              \
    } while (0)
    if (a.geom.mrows == 64) {
        if (a.stash) PDR_TC_LAUNCH(64, true); else PDR_TC_LAUNCH(64, false);
    } else {
        if (a.stash) PDR_TC_LAUNCH(128, true); else PDR_TC_LAUNCH(128, false);
    }
#undef PDR_TC_LAUNCH
    PDR_CUDA_CHECK(cudaGetLastError());
    return PDEREAD_OK;
}

template <int ACT, int NX, int NT>
int launch_tc_bwd(const TcBwdArgs& a, int grid, cudaStream_t stream) {
    const size_t smem = a.geom.smem_bytes;
    if (a.geom.mrows == 64) {
        const int rc = ensure_dynamic_smem((const void*)tc_jet_bwd_kernel<ACT, NX, NT, 64>, smem);
        if (rc != PDEREAD_OK) return rc;
        tc_jet_bwd_kernel<ACT, NX, NT, 64><<<grid, TC_THREADS, smem, stream>>>(a);
    } else {
        const int rc = ensure_dynamic_smem((const void*)tc_jet_bwd_kernel<ACT, NX, NT, 128>, smem);
        if (rc != PDEREAD_OK) return rc;
        tc_jet_bwd_kernel<ACT, NX, NT, 128><<<grid, TC_THREADS, smem, stream>>>(a);
    }
    PDR_CUDA_CHECK(cudaGetLastError());
    return PDEREAD_OK;
}

}  // namespace pdr
