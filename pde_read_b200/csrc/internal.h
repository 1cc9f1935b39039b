// Internal declarations shared by the translation units of libpderead_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pderead_b200.h"

namespace pdr {

constexpr int MAXL = PDEREAD_MAX_HIDDEN_LAYERS;
constexpr int MAXD = PDEREAD_MAX_INPUT_DIM;

// Thread mapping of the MLP-jet kernels: lane = collocation point, warp = chunk of TN neurons.
constexpr int TN = 10;          // output neurons per thread in the layer products
constexpr int TNP = 12;         // chunk length in the packed weights (keeps rows 16-byte aligned)
constexpr int MAX_WARPS = 12;   // warps per CTA (chunks beyond that are looped over)

// points per lane as a function of the jet length J = 1 + NX + NT
template <int J>
struct PointsPerLane {
    static constexpr int v = (J == 1) ? 4 : ((J <= 4) ? 2 : 1);
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
    const float* wt_packed;   // [L-1][W][NC*TNP], wt[l][k][c*TNP+n] = W_{l+1}[c*TN+n][k]
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
};

struct BwdArgs {
    NetDev net;
    GradLayout lay;
    const float* wn_packed;   // [L-1][W][NC*TNP], wn[l][i][c*TNP+n] = W_{l+1}[i][c*TN+n]
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

// geometry helpers shared by host code
inline int jet_len(int nx, int nt) { return 1 + nx + nt; }
inline int points_per_lane(int J) { return (J == 1) ? 4 : ((J <= 4) ? 2 : 1); }
inline int tile_points(int J) { return 32 * points_per_lane(J); }
inline int h_stride(int J) { return J * tile_points(J) + 4; }
size_t fwd_smem_bytes(int W, int din, int J, int nwarps);
size_t bwd_smem_bytes(int W, int L, int din, int J);
int block_warps(int W);

void set_last_cuda_error(cudaError_t e);
void stage_mark(cudaStream_t s);   // measurement hook, see pderead_debug_stage_events

// extraction.cu
int library_gram_launch(const float* dxn, const float* b, long long M, int n, int K, double* G, double* Atb,
                        double* btb, cudaStream_t s);

}  // namespace pdr

#define PDR_CUDA_CHECK(expr)                              \
    do {                                                  \
        cudaError_t _e = (expr);                          \
        if (_e != cudaSuccess) {                          \
            pdr::set_last_cuda_error(_e);                 \
            return PDEREAD_ERR_CUDA;                      \
        }                                                 \
    } while (0)
