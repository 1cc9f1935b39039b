// Dev probe (GPU): cost of fetching warp-uniform weights for the layer-product inner loop
//   (a) through the read-only global path as three uniform LDG.128 per k-step (what the kernels do),
//   (b) from the constant bank (LDC with a register index),
// each feeding 20 FFMA2 (J = 4, TN = 10) per k-step, 20 warps per SM, weights of a 5x50 network (48 KB packed).
// Prints ms per pass; results feed DESIGN.md "notes for the next round".
#include <cstdio>
#include <cuda_runtime.h>

constexpr int W = 50, NC = 5, TNP = 12, LAYERS = 4;
constexpr int PACK = LAYERS * W * NC * TNP;      // 12000 floats = 48 KB
__constant__ float c_w[PACK];

template <bool CONST>
__global__ void __launch_bounds__(160) probe(const float* __restrict__ gw, float* out, int tiles) {
    const int warp = threadIdx.x >> 5;
    float2 acc[4][5];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int n = 0; n < 5; ++n) acc[j][n] = make_float2(0.f, 0.f);
    const float h0 = (float)threadIdx.x * 1e-3f;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        for (int l = 0; l < LAYERS; ++l) {
            int base = (l * W * NC + warp) * TNP;
#pragma unroll 2
            for (int k = 0; k < W; ++k) {
                float4 b0, b1, b2;
                if (CONST) {
                    b0 = *reinterpret_cast<const float4*>(&c_w[base]);
                    b1 = *reinterpret_cast<const float4*>(&c_w[base + 4]);
                    b2 = *reinterpret_cast<const float4*>(&c_w[base + 8]);
                } else {
                    b0 = __ldg(reinterpret_cast<const float4*>(gw + base));
                    b1 = __ldg(reinterpret_cast<const float4*>(gw + base) + 1);
                    b2 = __ldg(reinterpret_cast<const float4*>(gw + base) + 2);
                }
                const float2 bw[5] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y),
                                      make_float2(b1.z, b1.w), make_float2(b2.x, b2.y)};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float hv = h0 + (float)(j + k);
                    const float2 a2 = make_float2(hv, hv);
#pragma unroll
                    for (int n = 0; n < 5; ++n) acc[j][n] = __ffma2_rn(a2, bw[n], acc[j][n]);
                }
                base += NC * TNP;
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int n = 0; n < 5; ++n) s += acc[j][n].x + acc[j][n].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float *gw, *out;
    cudaMalloc(&gw, PACK * sizeof(float));
    cudaMalloc(&out, 592 * 160 * sizeof(float));
    float* h = new float[PACK];
    for (int i = 0; i < PACK; ++i) h[i] = 1e-3f * (float)(i % 97);
    cudaMemcpy(gw, h, PACK * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(c_w, h, PACK * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int tiles = 3125;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode) probe<true><<<592, 160>>>(gw, out, tiles);
            else      probe<false><<<592, 160>>>(gw, out, tiles);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("%s rep %d: %.3f ms  (%.1f TFLOP/s)\n", mode ? "constant bank" : "LDG.128      ", rep, ms,
                   2.0 * tiles * LAYERS * W * 40.0 * 32 * NC / (ms * 1e-3) / 1e12);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("done: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
