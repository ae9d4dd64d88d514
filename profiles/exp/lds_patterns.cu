// Micro-benchmark: shared-memory load cost (cycles per warp instruction at saturation) for the access patterns the
// RoIAlign3D forward kernel can choose from.  nvcc -arch=sm_100a -O3 -o lds_patterns lds_patterns.cu && ./lds_patterns
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters) {
    __shared__ __align__(16) float s[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned base = (unsigned)__cvta_generic_to_shared(s);
    unsigned off;
    if (MODE == 0) off = lane * 4;                 // LDS.32 consecutive
    if (MODE == 1) off = (lane & 7) * 8;           // LDS.64, 8 words, 4-way broadcast
    if (MODE == 2) off = lane * 8;                 // LDS.64 distinct
    if (MODE == 3) off = 0;                        // LDS.128 uniform
    if (MODE == 4) off = (lane & 7) * 16;          // LDS.128, 8 x 16 B, 4-way broadcast
    if (MODE == 5) off = lane * 16;                // LDS.128 distinct
    if (MODE == 6) off = (lane & 3) * 8;           // LDS.64, 4 words, 8-way broadcast
    if (MODE == 7) off = 0;                        // LDS.32 uniform
    if (MODE == 8) off = 0;                        // LDS.64 uniform
    if (MODE == 9) off = (lane & 15) * 8;          // LDS.64, 16 words, 2-way broadcast
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    unsigned addr = base + off;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const unsigned a = addr + ((i + u) & 7) * 512;
            float x, y, z, w;
            if (MODE == 0 || MODE == 7) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a)); a0 += x; }
            else if (MODE == 1 || MODE == 2 || MODE == 6 || MODE == 8 || MODE == 9) { asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(a)); a0 += x; a1 += y; }
            else { asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a)); a0 += x; a1 += y; a2 += z; a3 += w; }
        }
    }
    if (a0 + a1 + a2 + a3 == 1234.5f) out[0] = a0;
}

template <int MODE> void run(const char* name) {
    float* d; cudaMalloc(&d, 4);
    const int iters = 2000, threads = 512, blocks = 148 * 2;
    k<MODE><<<blocks, threads>>>(d, 10);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<MODE><<<blocks, threads>>>(d, iters); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    // warp instructions per SM: 2 blocks * 16 warps * iters * 8
    const double instr_per_sm = 2.0 * 16 * iters * 8;
    const double cycles = ms * 1e-3 * 1.965e9;
    printf("%-40s %.3f ms  %.2f cycles per warp instruction (at 1965 MHz)\n", name, ms, cycles / instr_per_sm);
    cudaFree(d);
}

int main() {
    run<0>("LDS.32 consecutive");
    run<7>("LDS.32 uniform");
    run<8>("LDS.64 uniform");
    run<6>("LDS.64 4 words, 8-way broadcast");
    run<1>("LDS.64 8 words, 4-way broadcast");
    run<9>("LDS.64 16 words, 2-way broadcast");
    run<2>("LDS.64 distinct (256 B)");
    run<3>("LDS.128 uniform");
    run<4>("LDS.128 8 x 16 B, 4-way broadcast");
    run<5>("LDS.128 distinct (512 B)");
    return 0;
}
