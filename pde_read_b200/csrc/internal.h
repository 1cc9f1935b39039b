// Internal declarations shared by the translation units of libpderead_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pderead_b200.h"

namespace pdr {

constexpr int MAXL = PDEREAD_MAX_HIDDEN_LAYERS;
constexpr int MAXD = PDEREAD_MAX_INPUT_DIM;

// Thread mappings of the layer products in the MLP-jet kernels (ProdTile in mlp_jet_kernels.cuh): a warp owns a
// chunk of TN output neurons and the tile's points; the pair tile (J <= 4) splits the chunk into two halves of TH
// neurons over the lane parity, the lane tile (J >= 5) gives every lane one point and all TN neurons.
constexpr int TN = 10;          // output neurons per warp chunk
constexpr int TH = 5;           // pair tile: neurons per thread (half chunk)
constexpr int THP = 8;          // pair tile: half chunk padded to two 16-byte words in the packed weights
constexpr int TNP_PAIR = 2 * THP;   // packed floats per chunk and row, pair tile: [half][THP]
constexpr int TNP_LANE = 12;        // packed floats per chunk and row, lane tile (keeps rows 16-byte aligned)
constexpr int TNP = TNP_PAIR;       // workspace sizing: the larger of the two
#ifndef PDR_PROD_UNROLL
#define PDR_PROD_UNROLL 2           // unroll factor of the layer-product loop
#endif
#ifndef PDR_PAIR_MAXJ
#define PDR_PAIR_MAXJ 4             // longest jet that uses the pair tile (A/B builds: PDR_EXTRA_CFLAGS=-DPDR_PAIR_MAXJ=..)
#endif
inline bool use_pair_map(int J) { return J <= PDR_PAIR_MAXJ; }    // UsePairMap<J>
constexpr int MAX_WARPS = 16;   // warps per CTA of the forward kernel (chunks beyond that are looped over)
constexpr int BWD_WIDE_WARPS = 16;   // warps of the reverse kernel when only one CTA fits an SM

// points per lane as a function of the jet length J = 1 + NX + NT
#ifndef PDR_PPL_J1
#define PDR_PPL_J1 2
#endif
#ifndef PDR_PPL_J4
#define PDR_PPL_J4 1
#endif
// points per lane of forward-only (no stash) launches of the plain MLP, J = 1
constexpr int FWD_ONLY_PPL_J1 = 4;

template <int J>
struct PointsPerLane {
    static constexpr int v = (J == 1) ? PDR_PPL_J1 : ((J <= 4) ? PDR_PPL_J4 : 1);
};

// Register budget of the CUDA-core kernels.  They are latency-bound in the activation-jet code, so resident
// warps matter more than registers per thread: the launch bound is only a handle on ptxas' register cap,
// 65536 / max_threads (blocks are at most MAX_WARPS * 32 threads).  J <= 4: 96 registers -> 4 CTAs of 5 warps
// per SM at width 50 (tile = 32 points, 53 KB of shared memory); J >= 5: 128 (3 CTAs at width 50; wide networks
// whose tile fills the shared memory of an SM run one CTA of 16 warps instead, see bwd_block_warps).
// Measured on C2: 2 -> 4 CTAs/SM = 1.29x on the whole step.
template <int J>
struct RegCap {
    static constexpr int max_threads = (J <= 4) ? 640 : 512;
};

struct NetDev {
    int L;      // hidden layers
    int W;      // width
    int din;    // input dim
    int NC;     // ceil(W / TN)
    const float* w[MAXL + 1];
    const float* b[MAXL + 1];
    const float* ra[MAXL];
    const float* rb[MAXL];
};

// Offsets of every parameter tensor inside the flat "partial gradient" vector each CTA owns.
struct GradLayout {
    int off_w[MAXL + 1];
    int off_b[MAXL + 1];
    int off_ab[MAXL];   // 7 floats per layer: a[4] then b[3]
    int total;          // padded to a multiple of 4
};

struct FwdArgs {
    NetDev net;
    const float* wt_packed;   // [L-1][W][NC*WROW] transposed weights in the layout of the kernel's ProdTile (pack_weights_kernel)
    const float* in;          // [M][din]
    long long M;
    long long num_tiles;
    float* out0;              // J == 1: out [M];  else dxn [M][NX+1]
    float* out1;              // dtm [M] (J > 1, NT > 0; may be null)
    float* stash;             // pre-activations of hidden layers 1..L-1 (null unless STASH)
    // residual epilogue (J == 1 only): R = dtm_in - out
    const float* dtm_in;
    float* resid_out;
    float* seed_out;          // -seed_scale * R  = d(loss)/d(N output)
    const float* grad_resid;  // optional externally supplied dL/dR (then seed_out = -grad_resid)
    float seed_scale;
    double* sumsq;
    int aux;                  // mlp_plain_fwd_kernel: number of weight buffers in shared memory
    int aux2;                 // mlp_plain_fwd_kernel with a stash: points per lane of the stash tile (2: 64 points, 4: 128)
    float* hbuf;              // mlp_jet_fwd_kernel: activation buffers in the workspace ([grid][2][W][HS]) for networks whose
                              // tile does not fit the shared memory of an SM; null = shared memory
};

struct BwdArgs {
    NetDev net;
    GradLayout lay;
    const float* wn_packed;   // [L-1][W][NC*WROW] native weights in the layout of the kernel's ProdTile
    const float* in;          // [M][din]
    long long M;
    long long num_tiles;
    const float* stash;
    const float* g0;          // J == 1: grad_out [M]; else grad_dxn [M][NX+1] (may be null)
    const float* g1;          // grad_dtm [M] (may be null)
    float g0_scale, g1_scale;
    float* gin;               // [M][din] or null
    float* gpart;             // [gridDim.x][lay.total]
    int wg_mode;
    int accumulate;           // 0: this launch zeroes its partial-gradient vectors first
    int aux;                  // mlp_plain_bwd_kernel: number of weight buffers in shared memory
    float* hbuf;              // mlp_jet_bwd_kernel: Zb / Yb in the workspace ([grid][2][W][HS]) when they do not fit shared memory
};

// ---- tensor-core path (tc_jet_kernels.cuh) ----------------------------------------------------
struct TcGeom {
    int mrows;        // MMA M: 64 (width <= 64, two point sets in the lane halves) or 128
    int Kpad;         // width rounded up to 8
    int NB;           // neurons per TMEM lane block (16 lanes for M = 64, 32 for M = 128)
    int PC;           // points per chunk (per sub)
    int NCpad;        // MMA N = accumulator columns per chunk (PC * J rounded up)
    int TP;           // points per tile = subs * 2 chunks * PC
    int a_bytes;      // bytes of one weight image half (hi or lo)
    int a_rs;         // byte stride between 8-row groups of a weight image
    int slot_bytes;   // bytes of one operand slot half (hi or lo)
    int b_rs;         // byte stride between 8-column groups of an operand slot
    int sub_stride;   // byte stride between the slot sets of the two subs
    int nwbuf;        // weight buffers (1 or 2)
    int tmem_cols;    // TMEM columns allocated (power of two)
    size_t smem_bytes;
};

struct TcFwdArgs {
    NetDev net;
    TcGeom geom;
    const float* wimg;        // [L][hi|lo] weight images (tc_pack_weights_kernel)
    const float* in;          // [M][din]
    long long M;
    long long num_tiles;
    float* out0;
    float* out1;
    const float* dtm_in;
    float* resid_out;
    float* seed_out;
    const float* grad_resid;
    float seed_scale;
    double* sumsq;
};

struct Segment {
    float* dst;
    int off;
    int count;
};
struct SegmentTable {
    Segment seg[4 * MAXL + 2];
    int n;
};

// launchers; explicitly instantiated per (activation, NX, NT) in inst.cu
template <int ACT, int NX, int NT>
int launch_fwd(const FwdArgs& a, cudaStream_t stream);
template <int ACT, int NX, int NT>
int bwd_grid(int W, int L, int din, long long num_tiles, int* grid_out);
template <int ACT, int NX, int NT>
int launch_bwd_impl(const BwdArgs& a, int grid, cudaStream_t stream);

template <int ACT>
int launch_plain_fwd(const FwdArgs& a, cudaStream_t stream);   // register-tiled plain MLP (J = 1), mlp_plain_kernels.cuh
template <int ACT>
int plain_bwd_grid(int W, int L, int din, long long num_tiles, int* grid_out);
template <int ACT>
int launch_plain_bwd(const BwdArgs& a, int grid, cudaStream_t stream);

template <int ACT, int NX, int NT>
int launch_tc_fwd(const TcFwdArgs& a, cudaStream_t stream);

// geometry helpers shared by host code
inline int jet_len(int nx, int nt) { return 1 + nx + nt; }
inline int points_per_lane(int J) { return (J == 1) ? PDR_PPL_J1 : ((J <= 4) ? PDR_PPL_J4 : 1); }
inline int tile_points(int J) { return 32 * points_per_lane(J); }
inline int h_stride(int J) { return J * tile_points(J) + 4; }
size_t fwd_smem_bytes(int W, int din, int J, int nwarps);
size_t fwd_smem_bytes_tile(int W, int din, int J, int nwarps, int TP);   // for an explicit tile of TP points
size_t bwd_smem_bytes(int W, int L, int din, int J);
// Width ceiling: when the two W x HS activation buffers of a tile exceed the shared memory of an SM the kernels keep
// them in the workspace instead (one CTA per SM, L2-resident); these say when, and how much shared memory is left
bool fwd_spills(int W, int din, int J);
bool bwd_spills(int W, int L, int din, int J);
size_t fwd_hbuf_floats(int W, int J);     // per CTA
size_t bwd_hbuf_floats(int W, int J);
size_t fwd_smem_small(int W, int din, int J, int nwarps);
size_t bwd_smem_small(int W, int L, int din, int J);
bool bwd_stage_weights(int W, int L, int din, int J);
// forward kernel: bytes of the per-warp weight slices to add to `base_smem`, or 0 when staging is off for this shape
size_t fwd_wstage_bytes(int W, int L, int J, size_t base_smem);   // reverse kernel: stage the layer's weights in shared memory
int block_warps(int W);
int bwd_block_warps(int W, int L, int din, int J);

void set_last_cuda_error(cudaError_t e, const char* file, int line);
// cudaFuncAttributeMaxDynamicSharedMemorySize is one number per (kernel, device) for the whole process while one
// kernel serves several network shapes from several threads (autograd runs reverse calls on its own): this
// only ever raises it, so a shape seen earlier can always be launched again.
int ensure_dynamic_smem(const void* kernel, size_t smem);
int check_device();                // PDEREAD_OK on compute capability 10.x, else PDEREAD_ERR_NOT_BLACKWELL
// measurement hook, see pderead_debug_stage_events / pderead_debug_stage_tags: records the next caller-owned event
// after a launch, together with what was launched
enum StageTag { ST_OTHER = 0, ST_PACK = 1, ST_FWD_JET = 2, ST_FWD_PLAIN = 3, ST_BWD_PLAIN = 4, ST_BWD_JET = 5,
                ST_REDUCE = 6, ST_GRAM = 7, ST_RFE = 8, ST_OPTIM = 9, ST_BN = 10 };
void stage_mark(cudaStream_t s, int tag = ST_OTHER);

// extraction.cu
int library_gram_launch(const float* dxn, const float* b, long long M, int n, int K, double* G, double* Atb,
                        double* btb, cudaStream_t s);

}  // namespace pdr

#define PDR_CUDA_CHECK(expr)                              \
    do {                                                  \
        cudaError_t _e = (expr);                          \
        if (_e != cudaSuccess) {                          \
            pdr::set_last_cuda_error(_e, __FILE__, __LINE__); \
            return PDEREAD_ERR_CUDA;                      \
        }                                                 \
    } while (0)
