// Tensor-core (tcgen05 / TMEM) MLP-jet FORWARD kernel for sm_100a (forward-only calls: Evaluate_Derivatives,
// PDE_Residual without grad, Discovery_Testing, extraction).  Round 1 also carried a reverse kernel on this path; it
// lost to the CUDA-core reverse kernel on every shape (3xTF32 needs zbar in two operand layouts and y in one, each as
// hi | lo: 6 fp32 copies of a tile in shared memory leave room for 16-point chunks only) and was removed in round 2.
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
template <int ACT, int NX, int NT, int MROWS>
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
#define PDR_TC_LAUNCH(MR)                                                                                                \
    do {                                                                                                                 \
        const int rc_ = ensure_dynamic_smem((const void*)tc_jet_fwd_kernel<ACT, NX, NT, MR>, smem);                      \
        if (rc_ != PDEREAD_OK) return rc_;                                                                               \
        tc_jet_fwd_kernel<ACT, NX, NT, MR><<<(int)grid, TC_THREADS, smem, stream>>>(a);                                  \
    } while (0)
    if (a.geom.mrows == 64) PDR_TC_LAUNCH(64); else PDR_TC_LAUNCH(128);
#undef PDR_TC_LAUNCH
    PDR_CUDA_CHECK(cudaGetLastError());
    return PDEREAD_OK;
}

}  // namespace pdr
