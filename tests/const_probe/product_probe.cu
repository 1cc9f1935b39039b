// Dev probe (GPU): ceiling of the layer-product inner loop (J = 4, width 50, 4 hidden-hidden layers, 100 k points)
// under different operand sources / mappings, 4 CTAs of 5 warps (or 2 CTAs of 10 warps) per SM:
//   0  weights: 3 uniform LDG.128 per k-step; activations synthetic (registers)          -- arithmetic + weight fetch
//   1  as 0, activations from shared memory (one LDS.128 per k-step)                      -- what the kernels do
//   2  weights staged in shared memory per layer (uniform LDS.128), activations from shared memory
//   3  64-point tiles: 2 points per lane x 5 neurons per warp (2 LDG.128 + 2 LDS.128 per k-step), 10 warps per CTA
#include <cstdio>
#include <cuda_runtime.h>

constexpr int W = 50, LAYERS = 4, J = 4;

template <int MODE, int UNROLL = 2, bool SCALAR = false>
__global__ void __launch_bounds__(MODE == 3 ? 320 : 160) probe(const float* __restrict__ gw, float* out, int tiles) {
    constexpr int TN = (MODE == 3) ? 5 : 10, TNP = (MODE == 3) ? 8 : 12, PPL = (MODE == 3) ? 2 : 1, NC = W / TN;
    constexpr int TP = 32 * PPL, HS = J * TP + 4;
    extern __shared__ __align__(16) float sm[];
    float* H = sm;                                  // [W][HS]
    float* WS = sm + W * HS;                        // [W][NC*TNP] (MODE 2)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < W * HS; i += blockDim.x) H[i] = 1e-3f * (float)(i % 101);
    __syncthreads();
    float2 acc[PPL][J][TN / 2 + (TN % 2)];
#pragma unroll
    for (int q = 0; q < PPL; ++q)
#pragma unroll
        for (int j = 0; j < J; ++j)
#pragma unroll
            for (int n = 0; n < (TN + 1) / 2; ++n) acc[q][j][n] = make_float2(0.f, 0.f);
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        for (int l = 0; l < LAYERS; ++l) {
            const float* wl = gw + (size_t)l * W * NC * TNP;
            if (MODE == 2) {
                __syncthreads();
                for (int i = threadIdx.x; i < W * NC * TNP / 4; i += blockDim.x)
                    reinterpret_cast<float4*>(WS)[i] = __ldg(reinterpret_cast<const float4*>(wl) + i);
                __syncthreads();
            }
            const float* wk = (MODE == 2 ? WS : wl) + warp * TNP;
            const float* hk = H + lane * J;
#pragma unroll UNROLL
            for (int k = 0; k < W; ++k) {
                float4 b0, b1, b2 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (MODE == 2) {
                    b0 = *reinterpret_cast<const float4*>(wk);
                    b1 = *reinterpret_cast<const float4*>(wk + 4);
                    b2 = *reinterpret_cast<const float4*>(wk + 8);
                } else {
                    b0 = __ldg(reinterpret_cast<const float4*>(wk));
                    b1 = __ldg(reinterpret_cast<const float4*>(wk) + 1);
                    if (MODE != 3) b2 = __ldg(reinterpret_cast<const float4*>(wk) + 2);
                }
                const float2 bw[6] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y),
                                      make_float2(b1.z, b1.w), make_float2(b2.x, b2.y), make_float2(b2.z, b2.w)};
#pragma unroll
                for (int q = 0; q < PPL; ++q) {
                    float4 hv;
                    if (MODE == 0) hv = make_float4((float)k, (float)(k + 1), (float)(k + 2), (float)(k + 3));
                    else           hv = *reinterpret_cast<const float4*>(hk + q * 32 * J);
                    const float hj[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        const float2 a2 = make_float2(hj[j], hj[j]);
#pragma unroll
                        for (int n = 0; n < (TN + 1) / 2; ++n) {
                            if (SCALAR) {
                                acc[q][j][n].x = fmaf(a2.x, bw[n].x, acc[q][j][n].x);
                                acc[q][j][n].y = fmaf(a2.y, bw[n].y, acc[q][j][n].y);
                            } else {
                                acc[q][j][n] = __ffma2_rn(a2, bw[n], acc[q][j][n]);
                            }
                        }
                    }
                }
                wk += NC * TNP;
                hk += HS;
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < PPL; ++q)
#pragma unroll
        for (int j = 0; j < J; ++j)
#pragma unroll
            for (int n = 0; n < (TN + 1) / 2; ++n) s += acc[q][j][n].x + acc[q][j][n].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int UNROLL = 2, bool SCALAR = false>
void run(const float* gw, float* out, const char* name) {
    constexpr int PPL = (MODE == 3) ? 2 : 1, TP = 32 * PPL, HS = J * TP + 4;
    constexpr int threads = (MODE == 3) ? 320 : 160, per_sm = (MODE == 3) ? 2 : 4;
    const size_t smem = sizeof(float) * (W * HS * 2 + (MODE == 2 ? W * 60 : 0));     // two activation buffers as in the kernels
    cudaFuncSetAttribute(probe<MODE, UNROLL, SCALAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int tiles = 100000 / TP;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        probe<MODE, UNROLL, SCALAR><<<148 * per_sm, threads, smem>>>(gw, out, tiles);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    // useful FMAs: tiles * TP points * LAYERS * W * W * J (padding of the odd 5-neuron chunk not counted)
    printf("%-58s %.3f ms  %.1f TFLOP/s\n", name, best, 2.0 * tiles * TP * LAYERS * (double)W * W * J / (best * 1e-3) / 1e12);
}

int main() {
    float *gw, *out;
    const int PACK = LAYERS * W * 64;
    cudaMalloc(&gw, PACK * sizeof(float));
    cudaMalloc(&out, 148 * 4 * 320 * sizeof(float));
    float* h = new float[PACK];
    for (int i = 0; i < PACK; ++i) h[i] = 1e-3f * (float)(i % 97);
    cudaMemcpy(gw, h, PACK * sizeof(float), cudaMemcpyHostToDevice);
    run<0>(gw, out, "0 LDG weights, synthetic activations");
    run<1>(gw, out, "1 LDG weights, activations from shared memory");
    run<2>(gw, out, "2 weights staged in shared memory, activations from smem");
    run<3>(gw, out, "3 64-point tiles (2 points/lane x 5 neurons), LDG weights");
    run<1, 1>(gw, out, "1 with unroll 1");
    run<1, 4>(gw, out, "1 with unroll 4");
    run<1, 2, true>(gw, out, "1 with scalar FFMA instead of FFMA2");
    run<0, 2, true>(gw, out, "0 with scalar FFMA instead of FFMA2");
    cudaError_t e = cudaDeviceSynchronize();
    printf("done: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
