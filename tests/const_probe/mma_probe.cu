// Dev probe (GPU): throughput of the legacy warp-level tensor path (mma.sync.m16n8k8 tf32, SASS HMMA) on sm_100a with
// operands in registers -- the candidate for the weight-gradient product gW += zbar y^T of the reverse kernels, whose
// operands sit in shared memory as plain fp32 rows (any layout can be gathered into mma.sync fragments by LDS, whereas
// tcgen05 needs both operands in the 8 x 16-byte core-matrix layout, twice for hi / lo: that does not fit).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NACC>
__global__ void __launch_bounds__(256) probe(float* out, int iters) {
    float d[NACC][4];
    unsigned a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
    b[0] = __float_as_uint(0.5f); b[1] = __float_as_uint(0.25f);
#pragma unroll
    for (int n = 0; n < NACC; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) d[n][i] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int n = 0; n < NACC; ++n) mma_tf32(d[n], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int n = 0; n < NACC; ++n) s += d[n][0] + d[n][1] + d[n][2] + d[n][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
void run(float* out, int ctas_per_sm, const char* name) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4096;
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        probe<NACC><<<148 * ctas_per_sm, 256>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    const double flops = 2.0 * 16 * 8 * 8 * (double)NACC * iters * 148.0 * ctas_per_sm * 8;   // 8 warps per CTA
    printf("%-56s %.3f ms  %.1f TFLOP/s (tf32 mma.sync m16n8k8)\n", name, best, flops / (best * 1e-3) / 1e12);
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    run<4>(out, 1, "4 accumulators/warp, 8 warps/SM");
    run<8>(out, 1, "8 accumulators/warp, 8 warps/SM");
    run<8>(out, 2, "8 accumulators/warp, 16 warps/SM");
    run<8>(out, 4, "8 accumulators/warp, 32 warps/SM");
    run<16>(out, 2, "16 accumulators/warp, 16 warps/SM");
    cudaError_t e = cudaDeviceSynchronize();
    printf("done: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
