// Dev probe (GPU): shared-memory wavefronts per warp-wide LDS.128 / LDS.64 / LDS.32 for the address patterns the
// weight-gradient tiles can be built from.  Run under
//   ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum
// and divide.  Rows are 132 words apart (the kernels' row stride, 4 mod 32 banks).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ROW = 132;     // floats
constexpr int ITERS = 256;

template <int PAT, int VEC>
__global__ void lds_probe(float* out) {
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < 64 * ROW; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int row;
    switch (PAT) {
        case 0: row = 0; break;                                   // uniform
        case 1: row = lane / 16; break;                           // 2 runs
        case 2: row = lane / 8; break;                            // 4 runs (quarter aligned)
        case 3: row = lane < 13 ? 0 : (lane < 26 ? 1 : 2); break; // 3 unaligned runs
        case 4: row = lane % 8; break;                            // 8 distinct rows in every quarter
        case 5: row = -1; break;                                  // 32 consecutive vectors of one row
        case 6: row = lane % 13; break;                           // 13 distinct rows
        case 7: row = lane % 4; break;
        case 8: row = lane % 2; break;
        case 9: row = lane; break;                                // 32 distinct rows
        default: row = 0;
    }
    // patterns 10..: index of a 16-byte vector inside one contiguous block ("vector-interleaved" operand layout)
    int vecidx = -1;
    switch (PAT) {
        case 10: vecidx = lane % 8; break;
        case 11: vecidx = lane / 8; break;
        case 12: vecidx = lane / 4; break;
        case 13: vecidx = lane % 16; break;
        case 14: vecidx = lane / 2; break;
        case 15: vecidx = lane % 4; break;
        case 16: vecidx = lane / 16; break;
        case 17: vecidx = (lane % 8) * 2; break;      // 8 vectors, 32 bytes apart
        case 18: vecidx = (lane / 8) * 8; break;      // 4 runs, 128 bytes apart
        case 19: vecidx = (lane % 2) * 8 + lane / 8; break;
        default: break;
    }
    const float* base = (vecidx >= 0) ? sm + vecidx * 4 : ((row < 0) ? sm + lane * VEC : sm + row * ROW);
    if (vecidx >= 0) row = -2;
    float acc = 0.f;
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it) {
        const float* p = base + ((it & 7) * 4) * (row == -1 ? 32 : (row == -2 ? 64 : 1));
        const unsigned a = (unsigned)__cvta_generic_to_shared(p);
        if (VEC == 4) {
            float x, y, z, w;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a));
            acc += x + y + z + w;
        } else if (VEC == 2) {
            float x, y;
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(a));
            acc += x + y;
        } else {
            float x;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a));
            acc += x;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int PAT, int VEC>
void run(float* out) {
    lds_probe<PAT, VEC><<<8, 128, 64 * ROW * sizeof(float)>>>(out);
}

int main() {
    float* out;
    cudaMalloc(&out, 8 * 128 * sizeof(float));
#define ALL(V) run<0, V>(out); run<1, V>(out); run<2, V>(out); run<3, V>(out); run<4, V>(out); run<5, V>(out); \
               run<6, V>(out); run<7, V>(out); run<8, V>(out); run<9, V>(out);
    ALL(4) ALL(2) ALL(1)
    run<10, 4>(out); run<11, 4>(out); run<12, 4>(out); run<13, 4>(out); run<14, 4>(out); run<15, 4>(out);
    run<16, 4>(out); run<17, 4>(out); run<18, 4>(out); run<19, 4>(out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("done: %s (instructions per kernel: %d warps x %d)\n", cudaGetErrorString(e), 8 * 4, ITERS);
    return e != cudaSuccess;
}
