// Standalone probe for the tcgen05 building blocks the tensor-core kernels rely on (dev tool, run on a
// B200 through gpurun):  shared-memory matrix descriptors without swizzle (K-major and MN-major),
// the TF32 instruction descriptor, the TMEM lane mapping for M = 128 and M = 64 (incl. the second,
// interleaved M = 64 accumulator at lane offset 16), tcgen05.commit -> mbarrier, tcgen05.ld.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu && ./umma_probe
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../pde_read_b200/csrc/umma.cuh"

using namespace pdr::umma;

struct Case {
    int M, N, K;          // MMA shape (K total, multiple of 8)
    int b_mn_major;       // 1: B stored [k][n] with n contiguous (MN-major); 0: B stored [n][k] K-major
    int lane_off;         // D lane offset (0 or 16, M = 64 only)
    int swap_roles;       // 1: put the strides into the other descriptor field (hypothesis test)
    int split3;           // 1: 3xTF32 (hi/lo split of full fp32 operands)
};

// operand element (r, c) -> byte offset; r = row index of the "8 rows x 16 bytes" core matrix direction
// c = index along the 16-byte direction.  RS: stride between 8-row groups, CS: stride between 4-col groups
__host__ __device__ inline uint32_t core_off(int r, int c, uint32_t RS, uint32_t CS) {
    return (uint32_t)(r & 7) * 16u + (uint32_t)(c & 3) * 4u + (uint32_t)(r >> 3) * RS + (uint32_t)(c >> 2) * CS;
}

__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                    float* __restrict__ raw, Case cs) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int M = cs.M, N = cs.N, K = cs.K;
    // A: K-major [M][K]: core = 8 rows(m) x 4 k;  RS_A = (K/4)*128, CS_A = 128
    const uint32_t RS_A = (uint32_t)(K / 4) * 128u, CS_A = 128u;
    const uint32_t a_bytes = (uint32_t)(M / 8) * RS_A;
    // B MN-major: stored [k][n], core = 8 rows(k) x 4 n: RS_B = (N/4)*128, CS_B = 128
    // B K-major : stored [n][k], core = 8 rows(n) x 4 k: RS_B = (K/4)*128, CS_B = 128
    const uint32_t RS_B = (uint32_t)((cs.b_mn_major ? N : K) / 4) * 128u, CS_B = 128u;
    const uint32_t b_bytes = (uint32_t)((cs.b_mn_major ? K : N) / 8) * RS_B;
    unsigned char* sAh = smem;
    unsigned char* sAl = sAh + a_bytes;
    unsigned char* sBh = sAl + a_bytes;
    unsigned char* sBl = sBh + b_bytes;

    for (int idx = threadIdx.x; idx < M * K; idx += blockDim.x) {
        const int m = idx / K, k = idx - m * K;
        const float v = A[idx];
        const float hi = tf32_hi(v);
        *reinterpret_cast<float*>(sAh + core_off(m, k, RS_A, CS_A)) = cs.split3 ? hi : v;
        *reinterpret_cast<float*>(sAl + core_off(m, k, RS_A, CS_A)) = v - hi;
    }
    for (int idx = threadIdx.x; idx < K * N; idx += blockDim.x) {
        const int k = idx / N, n = idx - k * N;
        const float v = B[idx];   // B given as [K][N] row-major
        const float hi = tf32_hi(v);
        const uint32_t off = cs.b_mn_major ? core_off(k, n, RS_B, CS_B) : core_off(n, k, RS_B, CS_B);
        *reinterpret_cast<float*>(sBh + off) = cs.split3 ? hi : v;
        *reinterpret_cast<float*>(sBl + off) = v - hi;
    }
    fence_proxy_async();
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    if (threadIdx.x < 32) tmem_alloc(&tmem_base_s, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_tf32(M, N, /*a_mn_major=*/0, cs.b_mn_major);
        const uint32_t d_addr = tmem_base + ((uint32_t)cs.lane_off << 16);
        // K-major operand: LBO = stride between the two 16-byte K chunks (CS), SBO = 8-row group stride (RS)
        // MN-major operand: SBO = stride between 4-element MN groups (CS), LBO = stride between 8-row K groups (RS)
        uint32_t a_lbo = CS_A, a_sbo = RS_A;
        uint32_t b_lbo = cs.b_mn_major ? RS_B : CS_B, b_sbo = cs.b_mn_major ? CS_B : RS_B;
        if (cs.swap_roles) { uint32_t t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
        const uint32_t a_step = 2u * CS_A;                       // 8 k = two 16-byte chunks
        const uint32_t b_step = cs.b_mn_major ? RS_B : 2u * CS_B; // MN-major: one 8-row K group
        int first = 1;
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t ah = make_smem_desc(smem_u32(sAh) + ks * a_step, a_lbo, a_sbo);
            const uint64_t al = make_smem_desc(smem_u32(sAl) + ks * a_step, a_lbo, a_sbo);
            const uint64_t bh = make_smem_desc(smem_u32(sBh) + ks * b_step, b_lbo, b_sbo);
            const uint64_t bl = make_smem_desc(smem_u32(sBl) + ks * b_step, b_lbo, b_sbo);
            mma_tf32_ss(d_addr, ah, bh, idesc, !first);
            first = 0;
            if (cs.split3) {
                mma_tf32_ss(d_addr, al, bh, idesc, 1);
                mma_tf32_ss(d_addr, ah, bl, idesc, 1);
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        tmem_ld_x8(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        tmem_wait_ld();
        for (int t = 0; t < 8; ++t) raw[(size_t)(warp * 32 + lane) * N + c0 + t] = __uint_as_float(v[t]);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem_base, 256);
}

static float frand() { return (float)rand() / (float)RAND_MAX * 2.f - 1.f; }
static float to_tf32(float v) {
    uint32_t u;
    memcpy(&u, &v, 4);
    u &= 0xffffe000u;
    memcpy(&v, &u, 4);
    return v;
}

int main() {
    Case cases[] = {
        {128, 32, 16, 1, 0, 0, 0}, {128, 32, 16, 1, 0, 1, 0}, {128, 32, 16, 0, 0, 0, 0}, {128, 32, 16, 0, 0, 1, 0},
        {64, 32, 16, 1, 0, 0, 0},  {64, 32, 16, 1, 16, 0, 0}, {64, 64, 56, 1, 0, 0, 1},  {128, 96, 104, 1, 0, 0, 1},
        {128, 112, 128, 0, 0, 0, 1}, {64, 64, 64, 0, 0, 0, 1}, {64, 40, 56, 1, 0, 0, 1}, {128, 48, 104, 1, 0, 0, 1},
    };
    int ok_all = 1;
    for (const Case& cs : cases) {
        const int M = cs.M, N = cs.N, K = cs.K;
        std::vector<float> A((size_t)M * K), B((size_t)K * N), raw((size_t)128 * N, -777.f);
        for (auto& v : A) v = cs.split3 ? frand() : to_tf32(frand());
        for (auto& v : B) v = cs.split3 ? frand() : to_tf32(frand());
        float *dA, *dB, *dR;
        cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dR, raw.size() * 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dR, raw.data(), raw.size() * 4, cudaMemcpyHostToDevice);
        const size_t smem = 2 * (size_t)M * K * 4 + 2 * (size_t)K * N * 4 + 1024;
        cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe_kernel<<<1, 128, smem>>>(dA, dB, dR, cs);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(raw.data(), dR, raw.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int m = 0; m < M; ++m) {
            const int lane = (M == 128) ? m : ((m & 15) + 32 * (m >> 4) + cs.lane_off);
            for (int n = 0; n < N; ++n) {
                double s = 0;
                for (int k = 0; k < K; ++k) s += (double)A[(size_t)m * K + k] * (double)B[(size_t)k * N + n];
                maxerr = fmax(maxerr, fabs(s - (double)raw[(size_t)lane * N + n]));
                maxref = fmax(maxref, fabs(s));
            }
        }
        const double tol = cs.split3 ? 2e-6 : 1e-5;
        const int ok = maxerr <= tol * maxref;
        if (!ok && !cs.swap_roles) ok_all = 0;
        printf("M=%d N=%d K=%d b_mn_major=%d lane_off=%d swap=%d split3=%d : maxerr %.3e maxref %.3e rel %.3e %s\n", M, N, K,
               cs.b_mn_major, cs.lane_off, cs.swap_roles, cs.split3, maxerr, maxref, maxerr / maxref, ok ? "OK" : "MISMATCH");
        cudaFree(dA); cudaFree(dB); cudaFree(dR);
    }
    printf(ok_all ? "PROBE PASS\n" : "PROBE FAIL\n");
    return 0;
}
