/* pderead_b200.h -- C ABI of libpderead_b200.so (sm_100a).
 *
 * Drop-in boundary for the data-parallel hot path of punkduckable/PDE-READ.  The reference is
 * pure Python on PyTorch and has no FFI of its own; each entry point below states the reference
 * function (file:line under /root/reference) whose work it replaces.  The Python package
 * pde_read_b200/ keeps the reference's call signatures and binds these symbols with ctypes
 * (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *   - every pointer named d_* is DEVICE memory owned by the caller; the library never allocates
 *     or frees device memory.  No result depends on an earlier call; what the library does keep for
 *     the life of the process is launch bookkeeping (occupancy per kernel and shape, the dynamic
 *     shared-memory limit per kernel, the compute-capability check per device; mutex / atomics,
 *     safe from any host thread) and the pderead_debug_* switches (process-wide atomics meant for
 *     tests and benchmarks: do not flip them while other threads are inside the library);
 *   - every compute entry point returns PDEREAD_ERR_NOT_BLACKWELL on a device that is not
 *     compute capability 10.x (the kernels are built for sm_100a only);
 *   - all tensors are contiguous row-major float32 unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - every function returns 0 on success or a negative pderead_status; no exceptions cross.
 *   - workspace: query the *_workspace_bytes function, hand in a device buffer at least that
 *     large (16-byte aligned).  A larger buffer lets the library process more points per chunk.
 *   - network shapes: any width (the reference takes any Neurons_Per_Layer, Network.py:110-130),
 *     1..PDEREAD_MAX_HIDDEN_LAYERS hidden layers, input dimension <= PDEREAD_MAX_INPUT_DIM.  The
 *     kernels keep a tile's activations (2 x width x (32 J + 4) floats, J = 1 + n + m) in shared
 *     memory; when that exceeds the 227 KB of an SM (width > 148 at J = 6, > 127 at J = 7, > 214 at
 *     J = 4) the two buffers move to the workspace (one CTA per SM, L2-resident) -- same results,
 *     lower throughput.  PDEREAD_ERR_UNSUPPORTED_NET is left for shapes whose per-CTA parameter
 *     accumulators alone exceed shared memory (hidden layers x width beyond ~50 000).
 */
#ifndef PDEREAD_B200_H
#define PDEREAD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDEREAD_MAX_HIDDEN_LAYERS 16
#define PDEREAD_MAX_INPUT_DIM 8
#define PDEREAD_MAX_SPATIAL_ORDER 4
#define PDEREAD_MAX_TIME_ORDER 2
#define PDEREAD_MAX_LIBRARY_COLUMNS 256

typedef enum {
    PDEREAD_OK = 0,
    PDEREAD_ERR_BAD_ARGUMENT = -1,       /* null pointer, negative size, ... */
    PDEREAD_ERR_UNSUPPORTED_ORDER = -2,  /* derivative orders outside the instantiated set */
    PDEREAD_ERR_UNSUPPORTED_NET = -3,    /* width / depth / input dim / activation not supported */
    PDEREAD_ERR_WORKSPACE_TOO_SMALL = -4,
    PDEREAD_ERR_CUDA = -5,               /* a CUDA runtime call failed (see pderead_last_cuda_error) */
    PDEREAD_ERR_NOT_BLACKWELL = -6       /* device is not compute capability 10.x */
} pderead_status;

typedef enum {
    PDEREAD_ACT_TANH = 0,     /* torch.nn.Tanh              reference Network.py:158-160 */
    PDEREAD_ACT_SIN = 1,      /* Sin.forward                reference Network.py:60-67   */
    PDEREAD_ACT_RATIONAL = 2  /* Rational.forward (3,2)     reference Network.py:7-57    */
} pderead_activation;

/* One Neural_Network (reference Network.py:72-170): num_hidden_layers hidden Linear layers of
 * `width` neurons with an activation after each, then a Linear to one output.
 *   weight[i]: Layers[i].weight, row-major [out, in]  (i = 0 .. num_hidden_layers)
 *   bias[i]  : Layers[i].bias
 *   rat_a[i] : Activation_Functions[i].a [4], rat_b[i]: .b [3]  (Rational only, else NULL)
 * pderead_net_grad has the same shape and receives d(loss)/d(parameter). */
typedef struct {
    int32_t num_hidden_layers;
    int32_t width;
    int32_t input_dim;
    int32_t activation; /* pderead_activation */
    const float* weight[PDEREAD_MAX_HIDDEN_LAYERS + 1];
    const float* bias[PDEREAD_MAX_HIDDEN_LAYERS + 1];
    const float* rat_a[PDEREAD_MAX_HIDDEN_LAYERS];
    const float* rat_b[PDEREAD_MAX_HIDDEN_LAYERS];
} pderead_net;

typedef struct {
    float* weight[PDEREAD_MAX_HIDDEN_LAYERS + 1];
    float* bias[PDEREAD_MAX_HIDDEN_LAYERS + 1];
    float* rat_a[PDEREAD_MAX_HIDDEN_LAYERS];
    float* rat_b[PDEREAD_MAX_HIDDEN_LAYERS];
} pderead_net_grad;

const char* pderead_version(void);
const char* pderead_status_string(int status);
/* cudaError_t of the most recent failing CUDA call on this host thread (diagnostic only). */
int pderead_last_cuda_error(void);
/* "<source file>:<line>: <cudaGetErrorString>" of the CUDA call that failed last on this thread ("" if none). */
const char* pderead_last_cuda_error_where(void);

/* ---- Evaluate_Derivatives: reference PDE_Residual.py:9-147 ------------------------------------
 * d_coords [M,2] (col 0 = t, col 1 = x) -> d_dtm [M] = D_t^m U, d_dxn [M, n+1] = D_x^k U, k=0..n.
 * time_order m in 0..2 (m = 0: d_dtm is not computed and may be NULL -- the extraction path,
 * reference Extraction.py:305-341, never uses it); space_order n in 1..4 (or n = 0 with m = 0:
 * plain network evaluation U(coords), the Data_Loss forward, reference Loss_Functions.py:115). */
size_t pderead_sol_derivatives_workspace_bytes(const pderead_net* sol, int time_order, int space_order,
                                               int64_t M, int need_backward);
int pderead_sol_derivatives_forward(const pderead_net* sol, int time_order, int space_order,
                                    const float* d_coords, int64_t M, float* d_dtm, float* d_dxn,
                                    void* d_workspace, size_t workspace_bytes, void* stream);
/* Reverse pass of the above (what Loss.backward() does through the nested autograd graph,
 * reference Test_Train.py:99-100): given d_grad_dtm [M] (may be NULL = 0) and d_grad_dxn [M,n+1]
 * (may be NULL = 0), writes grads = beta*grads + sum_points(...) for every parameter of `sol`. */
int pderead_sol_derivatives_backward(const pderead_net* sol, int time_order, int space_order,
                                     const float* d_coords, int64_t M, const float* d_grad_dtm,
                                     const float* d_grad_dxn, const pderead_net_grad* grads, float beta,
                                     void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- Neural_Network.forward with scalar output: reference Network.py:174-202 ------------------
 * d_in [M, input_dim] -> d_out [M].  Used for PDE_NN(Dxn_U) (reference PDE_Residual.py:211). */
size_t pderead_mlp_workspace_bytes(const pderead_net* net, int64_t M, int need_backward);
int pderead_mlp_forward(const pderead_net* net, const float* d_in, int64_t M, float* d_out,
                        void* d_workspace, size_t workspace_bytes, void* stream);
/* d_grad_out [M] -> parameter grads (beta as above) and, if d_grad_in != NULL, d_grad_in [M,input_dim]. */
int pderead_mlp_backward(const pderead_net* net, const float* d_in, int64_t M, const float* d_grad_out,
                         float* d_grad_in, const pderead_net_grad* grads, float beta,
                         void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- PDE_Residual / Collocation_Loss: reference PDE_Residual.py:151-215, Loss_Functions.py:11-65
 * R = D_t^m U - N(U, D_x U, .., D_x^n U) at d_coords [M,2].
 *   d_resid [M]        : optional (NULL to skip)
 *   d_sum_sq (double*) : optional; the launch ADDS sum(R^2) over these M points to *d_sum_sq
 *                        (caller zeroes it; loss = *d_sum_sq / M_total). */
size_t pderead_residual_workspace_bytes(const pderead_net* sol, const pderead_net* pde, int time_order,
                                        int space_order, int64_t M, int need_backward);
int pderead_residual_forward(const pderead_net* sol, const pderead_net* pde, int time_order, int space_order,
                             const float* d_coords, int64_t M, float* d_resid, double* d_sum_sq,
                             void* d_workspace, size_t workspace_bytes, void* stream);
/* Collocation_Loss forward + Loss.backward() in one call (reference Test_Train.py:79-100 for the
 * collocation term): loss = sum(R^2)/M_total over the M local points (M_total >= M lets several
 * GPUs each take a shard; M_total = M on one GPU; M_total = 0 asks for the gradient of the UNSCALED
 * sum(R^2), which the sharded host path divides by the all-reduced point count).  Adds sum(R^2) to
 * *d_sum_sq and writes
 * grads = beta*grads + d(sum(R^2)/M_total)/d(parameter) for both networks.
 * If d_grad_resid != NULL it is used as d(loss)/dR [M] instead of 2R/M_total (PDE_Residual with an
 * arbitrary downstream loss). */
int pderead_collocation_loss_grad(const pderead_net* sol, const pderead_net* pde, int time_order,
                                  int space_order, const float* d_coords, int64_t M, int64_t M_total,
                                  const float* d_grad_resid, float* d_resid, double* d_sum_sq,
                                  const pderead_net_grad* sol_grads, const pderead_net_grad* pde_grads,
                                  float beta, void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- Extraction: reference Extraction.py:178-572 ----------------------------------------------
 * Library columns are the monomials of (U, D_x U, .., D_x^n U) of degree <= poly_degree in the
 * reference's order (Extraction.py:349-372); C = number of columns = binom(n+1+K, K). */
int pderead_library_num_columns(int space_order, int poly_degree);
/* multi-index table: h_indices [C, poly_degree] int32, padded with -1; row 0 is the constant. */
int pderead_library_multi_indices(int space_order, int poly_degree, int32_t* h_indices);
/* dense library (Generate_Library's second return value): d_dxn [M,n+1] -> d_library [M,C]. */
int pderead_library_dense(const float* d_dxn, int64_t M, int space_order, int poly_degree,
                          float* d_library, void* stream);
/* Fused library + Gram: accumulates, in float64, G += A^T A [C,C], Atb += A^T b [C], btb += b^T b
 * over M rows without materialising A.  Monomials are formed in float32 with the reference's
 * left-to-right product order, then widened.  Outputs are ADDED to (caller zeroes them). */
size_t pderead_library_gram_workspace_bytes(int space_order, int poly_degree);
int pderead_library_gram(const float* d_dxn, const float* d_b, int64_t M, int space_order, int poly_degree,
                         double* d_G, double* d_Atb, double* d_btb, void* d_workspace,
                         size_t workspace_bytes, void* stream);
/* Gram of an explicit matrix (for Recursive_Feature_Elimination(A, b) called with a dense A). */
int pderead_dense_gram(const float* d_A, const float* d_b, int64_t M, int C, double* d_G, double* d_Atb,
                       double* d_btb, void* d_workspace, size_t workspace_bytes, void* stream);
/* Coords -> Gram in one call: U jets (no t-series), b = N(jets), fused library+Gram. */
size_t pderead_extraction_gram_workspace_bytes(const pderead_net* sol, const pderead_net* pde, int space_order,
                                               int poly_degree, int64_t M);
int pderead_extraction_gram(const pderead_net* sol, const pderead_net* pde, int space_order, int poly_degree,
                            const float* d_coords, int64_t M, double* d_G, double* d_Atb, double* d_btb,
                            void* d_workspace, size_t workspace_bytes, void* stream);
/* Recursive_Feature_Elimination + Rank_Candidate_Solutions on the Gram system, on device
 * (reference Extraction.py:382-501 and 505-572).  Inputs d_G [C,C], d_Atb [C], d_btb [1] (float64).
 *   d_X [C,C] float32        : column j = j-th candidate (C-j non-zeros), un-scaled
 *   d_residual [C+1] float32 : ||b - A x_j||_2, last = ||b||_2
 *   d_order [C] int32        : feature eliminated at step j
 *   d_rank [C] int32, d_change [C] float32 : Rank_Candidate_Solutions' permutation (ascending
 *                              relative residual jump; last = most likely PDE) and sorted jumps. */
size_t pderead_rfe_workspace_bytes(int C);
int pderead_rfe_gram(const double* d_G, const double* d_Atb, const double* d_btb, int C, float* d_X,
                     float* d_residual, int32_t* d_order, int32_t* d_rank, float* d_change,
                     void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- Generate_Points: reference Points.py:7-62 --------------------------------------------------
 * Uniform points in the box h_bounds [dim,2] (host, float64), Philox-4x32-10 counter stream
 * (seed, offset + point index): distribution-equal, not bit-equal, to random.uniform. */
int pderead_generate_points(const double* h_bounds, int dim, int64_t num_points, uint64_t seed,
                            uint64_t offset, float* d_points, void* stream);

/* ---- input normalisation of the PDE network -------------------------------------------------
 * reference Network.py:100-104,194-195: with "PDE Network - Normalize Inputs" the PDE network applies
 * torch.nn.BatchNorm1d to its inputs [U, D_x U, ..]: the one operation of the path that couples the points.
 * Two phases: (1) pderead_input_moments accumulates {count, sum_d, sum_d of squares} in float64 (the host sums the
 * 1 + 2 din numbers over ranks when points are sharded); (2) pderead_bn_fold turns statistics + gamma/beta into
 * folded first-layer parameters (the normalisation is affine in the inputs), after which the ordinary MLP entry
 * points run on a pderead_net whose weight[0] / bias[0] point at the folded copies.  training != 0: batch statistics
 * (biased variance), running statistics updated with `momentum` (unbiased variance) and *d_num_batches += 1, as
 * BatchNorm1d does; training == 0: running statistics.  d_stats receives [mean (din) | rstd (din)].
 * Reverse: pderead_mlp_backward on the folded network yields gW0', gb0' and the input gradient at fixed statistics;
 * pderead_bn_grad_sums -> {sum_i gW0'[i][d] W0[i][d], sum_i gb0'[i] W0[i][d]} (sum these over ranks for the
 * `global` argument); pderead_bn_backward rewrites gW0' -> gW0 in place (gb0 = gb0'), writes the gradients of
 * gamma / beta (from the local sums) and the seeds a_d, b_d (from the global sums); pderead_bn_input_grad adds the
 * batch-statistics term a_d + b_d x_pd to the input gradient. */
int pderead_input_moments(const float* d_x, int64_t M, int din, double* d_moments, void* stream);
int pderead_bn_fold(const pderead_net* net, const float* d_gamma, const float* d_beta, const double* d_moments,
                    int training, float momentum, float eps, float* d_running_mean, float* d_running_var,
                    int64_t* d_num_batches, float* d_w0_folded, float* d_b0_folded, float* d_stats, void* stream);
int pderead_bn_grad_sums(const pderead_net* net, const float* d_gw0_folded, const float* d_gb0_folded, double* d_sums,
                         void* stream);
int pderead_bn_backward(const pderead_net* net, const float* d_gamma, const float* d_beta, const float* d_stats,
                        const double* d_moments, const double* d_sums_local, const double* d_sums_global, int training,
                        float* d_gw0, const float* d_gb0, float* d_ggamma, float* d_gbeta, float* d_seed_ab,
                        void* stream);
int pderead_bn_input_grad(float* d_gin, const float* d_x, int64_t M, int din, const float* d_seed_ab, void* stream);

/* ---- optimiser step -------------------------------------------------------------------
 * reference main.py:63-71 builds torch.optim.Adam / SGD(momentum 0.9, nesterov) over the ~50 parameter tensors
 * of both networks and Test_Train.py:105 steps it.  Here all parameters, gradients and optimiser state live in
 * flat float32 buffers of n elements with one common layout (every tensor 16-byte aligned, padding zero) and one
 * launch updates them, in torch's operation order.  d_state2 = {step count (float), internal ticket (zero it
 * once)}; the kernel advances the count itself, so the launch can sit inside a CUDA graph.  grad_scale multiplies
 * the gradients first (1 for a plain step). */
int pderead_adam_step(float* d_params, const float* d_grads, float* d_exp_avg, float* d_exp_avg_sq, int64_t n, float lr,
                      float beta1, float beta2, float eps, float weight_decay, float grad_scale, float* d_state2,
                      void* stream);
int pderead_sgd_step(float* d_params, const float* d_grads, float* d_momentum_buf, int64_t n, float lr, float momentum,
                     int nesterov, float weight_decay, float grad_scale, float* d_state2, void* stream);

/* ---- sharded points (one process per GPU): the cross-rank part of Collocation_Loss ----------------
 * mean(R^2) (reference Loss_Functions.py:65) and its gradient are sums over points, so every rank works on
 * its slice (pderead_collocation_loss_grad with M_total = 0: unscaled sums) and ONE float32 SUM all-reduce of
 * [gradients | tail4] finishes the step.  pack writes tail4 = {sum_sq hi, lo, M_local >> 12, M_local & 4095}
 * (exact in float32 sums); unpack reads the reduced tail: d_out3 = {sum(R^2) over all ranks, total point
 * count, scale / count} -- the factor the summed gradients are multiplied with (scale = 1 for the mean). */
int pderead_loss_tail_pack(const double* d_sum_sq, int64_t M_local, float* d_tail4, void* stream);
int pderead_loss_tail_unpack(const float* d_tail4, int64_t scale, double* d_out3, void* stream);

/* ---- the per-step exchange over NVLink peer memory (one box, one process per GPU) ----------------------------
 * Replaces the NCCL all-reduce of the packed float32 buffer [dU | dN | sum(R^2) | count] that ends a sharded
 * Collocation_Loss step (reference semantics: the mean of Loss_Functions.py:65 and the gradient sums of
 * Loss.backward(), Test_Train.py:99-100, over ALL points).  Every rank writes its buffer into a slot of every
 * peer's receive area (remote stores), raises a flag there, then sums the slots of its own area in rank order: one
 * kernel on the caller's stream, bit-identical results on all ranks, no host synchronisation.
 *   pderead_peer_alloc      cudaMalloc's this rank's receive area for `world` ranks and buffers of up to cap_floats
 *                           floats, zeroes it and returns its 64-byte CUDA IPC handle (the one place where the library
 *                           allocates device memory: the handle must cover exactly one allocation);
 *   pderead_peer_open       maps a peer's area from its handle (peer access is enabled on demand);
 *   pderead_peer_allreduce  d_buf[0..n) <- sum over ranks, in place; n % 4 == 0, d_buf 16-byte aligned;
 *                           d_peer_areas = DEVICE array of `world` area pointers (own area at index rank).  Collective:
 *                           every rank must call it the same number of times with the same n.
 */
size_t pderead_peer_area_bytes(int world, int cap_floats);
int pderead_peer_alloc(int world, int cap_floats, void** d_area, unsigned char* handle64);
int pderead_peer_open(const unsigned char* handle64, void** d_area);
int pderead_peer_close(void* d_area);
int pderead_peer_free(void* d_area);
int pderead_peer_allreduce(float* d_buf, int64_t n, void* const* d_peer_areas, void* d_local_area, int rank, int world,
                           int cap_floats, void* stream);

/* ---- measurement helper ---------------------------------------------------------------------------
 * Runs a register-resident FFMA loop on every SM (`iters` dependent FMA chains per thread) so that
 * bench.py can measure the FP32 FMA peak that bounds this path.  variant 0: scalar FFMA,
 * variant 1: packed fma.rn.f32x2, variant 2: float64 DFMA (the bound of the Gram kernel).
 * *h_flops receives the FLOP count of the launch. */
int pderead_fma_peak_probe(int variant, int iters, float* d_sink, double* h_flops, void* stream);
/* Hand the library `count` caller-created cudaEvent_t (host array).  Every kernel the following
 * calls on this host thread launch is followed by cudaEventRecord(events[i++], stream) until the
 * array is used up; pass NULL to stop.  Lets bench.py time individual kernels inside a normal step
 * without synchronising.  pderead_debug_stage_events_recorded() returns how many were recorded. */
int pderead_debug_stage_events(void* const* events, int count);
int pderead_debug_stage_events_recorded(void);
/* What each recorded event followed: 0 other, 1 weight packing, 2 jet forward (U), 3 plain-MLP forward (N),
 * 4 plain-MLP reverse, 5 jet reverse, 6 partial-gradient reduction, 7 library + Gram, 8 RFE + ranking,
 * 9 optimiser step, 10 input-normalisation helpers.  Fills h_tags[0..return value). */
int pderead_debug_stage_tags(int32_t* h_tags, int capacity);
/* Plain-MLP forward launches (jet length 1: the PDE network, Data_Loss) use the register-tiled kernel with the
 * layer's weights staged in shared memory (on, default) or the lane = point kernel (off; also PDEREAD_PLAIN_MLP=0).
 * A measurement switch; both produce the same numbers up to summation order.  Returns the previous setting. */
int pderead_debug_set_plain_mlp(int on);
/* Tensor-core policy of the forward-only jet calls (tcgen05, 3xTF32; widths 8..128, depth >= 2):
 *   0 off   - CUDA-core kernels everywhere;
 *   1 auto  - default: forward-only calls of networks wider than 64 (>= 3 hidden layers) use the tensor cores;
 *   2 force - the tensor-core forward for every forward-only call whose shape it supports.
 * The gradient path (forward with stash + reverse) always runs on the CUDA cores (fp32-exact accumulation).
 * Also selectable with the environment variable PDEREAD_TC=off|auto|force.  Returns the previous mode. */
int pderead_debug_set_tensor_cores(int mode);

#ifdef __cplusplus
}
#endif
#endif /* PDEREAD_B200_H */
