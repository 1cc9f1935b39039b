// Dev probe (GPU): register-tiled layer product for the MLP-jet kernels.  Mapping under test: warp = group of 8 output
// neurons, lane = PP points (all J jet coefficients of a point stay in one thread), weights of the layer staged in shared
// memory as [k][neuron] (uniform LDS.128), activations [k][point][jet] (one LDS.128 per point at J = 4).
// Per k-step and thread: 2 + PP LDS.128 feed 8 * PP * J FMAs (PP = 2, J = 4: 4 loads per 32 FFMA2; the lane = point /
// 10-neuron mapping of round 1 has 4 loads per 20 FFMA2, three of them uniform LDG.128 that cost ~4 wavefronts each).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int W = 50, LAYERS = 4;

template <int J, int PP, int NG, int CTAS>
__global__ void __launch_bounds__(32 * ((W + NG - 1) / NG), CTAS) probe(const float* __restrict__ gw, float* out, int tiles) {
    constexpr int NGR = (W + NG - 1) / NG, WP = NGR * NG;       // neuron groups, padded width
    constexpr int TP = 32 * PP, HS = J * TP + 4;
    extern __shared__ __align__(16) float sm[];
    float* H = sm;                                  // [W][HS]
    float* WS = sm + W * HS;                        // [W][WP]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < W * HS; i += blockDim.x) H[i] = 1e-3f * (float)(i % 101);
    __syncthreads();
    float2 acc[PP * J][NG / 2];
#pragma unroll
    for (int r = 0; r < PP * J; ++r)
#pragma unroll
        for (int n = 0; n < NG / 2; ++n) acc[r][n] = make_float2(0.f, 0.f);
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        for (int l = 0; l < LAYERS; ++l) {
            const float* wl = gw + (size_t)l * W * WP;
            __syncthreads();
            for (int i = threadIdx.x; i < W * WP / 4; i += blockDim.x)
                reinterpret_cast<float4*>(WS)[i] = __ldg(reinterpret_cast<const float4*>(wl) + i);
            __syncthreads();
            const float* wk = WS + warp * NG;
            const float* hk = H + lane * J;          // [k][q][lane][jet]: every LDS.128 of a warp is contiguous
#pragma unroll 2
            for (int k = 0; k < W; ++k) {
                float wv[NG];
#pragma unroll
                for (int n = 0; n < NG; n += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(wk + n);
                    wv[n] = v.x; wv[n + 1] = v.y; wv[n + 2] = v.z; wv[n + 3] = v.w;
                }
                float hv[PP * J];
#pragma unroll
                for (int r = 0; r < PP * J; r += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(hk + (r / J) * 32 * J + (r % J));
                    hv[r] = v.x; hv[r + 1] = v.y; hv[r + 2] = v.z; hv[r + 3] = v.w;
                }
#pragma unroll
                for (int r = 0; r < PP * J; ++r) {
                    const float2 a2 = make_float2(hv[r], hv[r]);
#pragma unroll
                    for (int n = 0; n < NG / 2; ++n)
                        acc[r][n] = __ffma2_rn(a2, make_float2(wv[2 * n], wv[2 * n + 1]), acc[r][n]);
                }
                wk += WP;
                hk += HS;
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < PP * J; ++r)
#pragma unroll
        for (int n = 0; n < NG / 2; ++n) s += acc[r][n].x + acc[r][n].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// operands in registers only: what the FMA pipe sustains with three distinct register-pair operands per FFMA2
template <int NA, int NB>
__global__ void __launch_bounds__(256) regprobe(float* out, int iters, float seed) {
    float2 acc[NA][NB];
    float a[NA];
    float2 b[NB];
#pragma unroll
    for (int i = 0; i < NA; ++i) a[i] = seed + i + threadIdx.x * 1e-3f;
#pragma unroll
    for (int n = 0; n < NB; ++n) b[n] = make_float2(seed * n, seed * (n + 1));
#pragma unroll
    for (int i = 0; i < NA; ++i)
#pragma unroll
        for (int n = 0; n < NB; ++n) acc[i][n] = make_float2(0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const float2 a2 = make_float2(a[i], a[i]);
#pragma unroll
            for (int n = 0; n < NB; ++n) acc[i][n] = __ffma2_rn(a2, b[n], acc[i][n]);
        }
#pragma unroll
        for (int i = 0; i < NA; ++i) a[i] += 1e-7f;      // keeps the loop from being hoisted
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NA; ++i)
#pragma unroll
        for (int n = 0; n < NB; ++n) s += acc[i][n].x + acc[i][n].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NA, int NB>
void runreg(float* out, const char* name) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e9f;
    const int iters = 4096;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        regprobe<NA, NB><<<148 * 2, 256>>>(out, iters, 1e-3f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    printf("%-64s %.3f ms  %.1f TFLOP/s\n", name, best, 2.0 * 2 * NA * NB * (double)iters * 148 * 2 * 256 / (best * 1e-3) / 1e12);
}

template <int J, int PP, int NG, int CTAS>
void run(const float* gw, float* out, const char* name) {
    constexpr int NGR = (W + NG - 1) / NG, WP = NGR * NG, TP = 32 * PP, HS = J * TP + 4;
    constexpr int threads = 32 * NGR;
    const size_t smem = sizeof(float) * (W * HS * 2 + W * WP);     // two activation buffers as in the kernels
    cudaFuncSetAttribute(probe<J, PP, NG, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe<J, PP, NG, CTAS>, threads, smem);
    const int tiles = 100000 / TP;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        probe<J, PP, NG, CTAS><<<148 * (occ > 0 ? occ : 1), threads, smem>>>(gw, out, tiles);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    printf("%-64s occ %d x %d thr, smem %zu: %.3f ms  %.1f TFLOP/s\n", name, occ, threads, smem, best,
           2.0 * tiles * TP * LAYERS * (double)W * W * J / (best * 1e-3) / 1e12);
}

int main() {
    float *gw, *out;
    const int PACK = LAYERS * W * 64;
    cudaMalloc(&gw, PACK * sizeof(float));
    cudaMalloc(&out, 148 * 8 * 512 * sizeof(float));
    float* h = new float[PACK];
    for (int i = 0; i < PACK; ++i) h[i] = 1e-3f * (float)(i % 97);
    cudaMemcpy(gw, h, PACK * sizeof(float), cudaMemcpyHostToDevice);
    runreg<8, 4>(out, "registers only: 8 x 4 FFMA2 (+8 FADD) per iteration, 16 warps/SM");
    runreg<4, 5>(out, "registers only: 4 x 5 FFMA2 (+4 FADD) per iteration, 16 warps/SM");
    run<4, 2, 8, 2>(gw, out, "J=4: 8 neurons x 2 points/lane, 7 warps");
    run<4, 2, 10, 2>(gw, out, "J=4: 10 neurons x 2 points/lane, 5 warps");
    run<4, 1, 8, 4>(gw, out, "J=4: 8 neurons x 1 point/lane, 7 warps");
    run<4, 2, 4, 2>(gw, out, "J=4: 4 neurons x 2 points/lane, 13 warps");
    run<4, 1, 12, 4>(gw, out, "J=4: 12 neurons x 1 point/lane, 5 warps");
    run<4, 1, 16, 4>(gw, out, "J=4: 16 neurons x 1 point/lane, 4 warps");
    run<4, 2, 12, 2>(gw, out, "J=4: 12 neurons x 2 points/lane, 5 warps");
    run<1, 8, 8, 2>(gw, out, "J=1: 8 neurons x 8 points/lane, 7 warps");
    run<1, 4, 8, 4>(gw, out, "J=1: 8 neurons x 4 points/lane, 7 warps");
    run<1, 4, 16, 4>(gw, out, "J=1: 16 neurons x 4 points/lane, 4 warps");
    cudaError_t e = cudaDeviceSynchronize();
    printf("done: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
