// Microbenchmark: issue rate of the minimum instructions a reducing epilogue can use (per SM sub-partition).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/minmax_rate tools/minmax_rate.cu
// Each thread keeps 8 independent accumulators and folds fresh values (xorshift, so nothing constant-folds)
// in an unrolled loop; reports warp-instructions per clock per SM for 4 / 8 / 16 warps on one SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

enum { FMIN2, FMIN3, IMIN2, IMIN3, HMIN2X2, FSETSEL, FFMA, UMIN2 };

template <int OP>
__device__ __forceinline__ void fold(uint32_t &acc, uint32_t a, uint32_t b) {
    if (OP == FMIN2) {
        asm volatile("min.f32 %0, %0, %1;" : "+r"(acc) : "r"(a));
    } else if (OP == FMIN3) {
        asm volatile("min.f32 %0, %0, %1, %2;" : "+r"(acc) : "r"(a), "r"(b));
    } else if (OP == IMIN2) {
        asm volatile("min.s32 %0, %0, %1;" : "+r"(acc) : "r"(a));
    } else if (OP == UMIN2) {
        asm volatile("min.u32 %0, %0, %1;" : "+r"(acc) : "r"(a));
    } else if (OP == IMIN3) {
        acc = (uint32_t)__vimin3_s32((int)acc, (int)a, (int)b);
    } else if (OP == HMIN2X2) {
        asm volatile("min.f16x2 %0, %0, %1;" : "+r"(acc) : "r"(a));
    } else if (OP == FSETSEL) {
        asm volatile("{.reg .pred p; setp.lt.f32 p, %1, %0; selp.b32 %0, %1, %0, p;}" : "+r"(acc) : "r"(a));
    } else if (OP == FFMA) {
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(acc) : "r"(a), "r"(b));
    }
}

template <int OP>
__global__ void __launch_bounds__(512, 1) rate_kernel(int iters, long long *cycles, uint32_t *sink, uint32_t seed) {
    uint32_t acc[8], v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        acc[k] = 0x7f000000u;
        v[k] = seed * (threadIdx.x + 1) * (k + 3);
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < 8; ++k) fold<OP>(acc[k], v[k], v[(k + 1) & 7]);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] ^= acc[(k + 3) & 7] >> 3;   // 8 extra ALU ops per 64 folds
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    uint32_t x = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) x ^= acc[k];
    if (x == 0x12345u) sink[0] = x;
}

template <int OP>
static void run(const char *name, long long *dcy, uint32_t *dsink) {
    const int iters = 4000;
    for (int warps : {4, 8, 16}) {
        rate_kernel<OP><<<1, warps * 32>>>(iters, dcy, dsink, 12345u);
        cudaDeviceSynchronize();
        rate_kernel<OP><<<1, warps * 32>>>(iters, dcy, dsink, 12345u);
        cudaError_t e = cudaDeviceSynchronize();
        long long cy;
        cudaMemcpy(&cy, dcy, sizeof(cy), cudaMemcpyDeviceToHost);
        const double winstr = (double)warps * iters * 64.0;
        printf("%-28s warps %2d : %9lld cycles  %.3f folds (warp-instr)/clk/SM = %.1f lanes/clk/SM  (%s)\n", name, warps,
               cy, winstr / cy, 32.0 * winstr / cy, cudaGetErrorString(e));
    }
}

int main() {
    long long *dcy;
    uint32_t *dsink;
    cudaMalloc(&dcy, 8);
    cudaMalloc(&dsink, 4);
    run<FFMA>("fma.rn.f32 (reference)", dcy, dsink);
    run<FMIN2>("min.f32 (2-input)", dcy, dsink);
    run<FMIN3>("min.f32 (3-input)", dcy, dsink);
    run<IMIN2>("min.s32", dcy, dsink);
    run<UMIN2>("min.u32", dcy, dsink);
    run<IMIN3>("__vimin3_s32", dcy, dsink);
    run<HMIN2X2>("min.f16x2", dcy, dsink);
    run<FSETSEL>("setp.lt.f32 + selp", dcy, dsink);
    return 0;
}
