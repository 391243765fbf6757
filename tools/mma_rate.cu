// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16, M=128) issued back to back by one
// thread, for several shapes / accumulator patterns.  No TMA, no epilogue: operands are whatever
// is in shared memory.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_rate tools/mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
    return (uint64_t)((a >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

// mode: 0 SS same accumulator, 1 SS two alternating accumulators, 2 TS same accumulator,
//       3 SS four alternating accumulators
__device__ int g_random_fill = 0;
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long *cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *sm = (uint8_t *)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += 128) {
        uint32_t v = 0x3c003c00u;
        if (g_random_fill) {  // pseudo-random bf16 pairs in (-1,1): realistic switching activity
            uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
            h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
            v = (h & 0x807f807fu) | 0x3c003c00u | ((h >> 3) & 0x03000300u);
        }
        ((uint32_t *)sm)[i] = v;
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    if (warp == 1) {
        const uint64_t a = desc_sw128(smem_u32(sm)), b = desc_sw128(smem_u32(sm + 16384));
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int i = 0; i < iters; i += 4) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t d = tb;
                    if (MODE == 1) d = tb + ((i + k) & 1) * N;
                    if (MODE == 3) d = tb + ((i + k) & 3) * 128;
                    if (MODE == 2) mma_ts(d, tb + 256 + 8 * k, b + 2 * k, idesc, 1);
                    else mma_ss(d, a + 2 * k, b + 2 * k, idesc, 1);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
            t1 = clock64();
            cycles[blockIdx.x] = t1 - t0;
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

// The issue pattern of a pipelined GEMM: BATCH MMAs, tcgen05.commit, then the bookkeeping that
// precedes the next batch (mbarrier probe that succeeds at once, tcgen05 fence, elect).
template <int BATCH>
__global__ void __launch_bounds__(128, 1) batch_kernel(int iters, long long *cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, bar2, ready;
    __shared__ uint32_t tmem_base_s;
    uint8_t *sm = (uint8_t *)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (16384 * 2) / 4; i += 128) ((uint32_t *)sm)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ready)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | (8u << 24);
    if (warp == 1) {
        const uint64_t a = desc_sw128(smem_u32(sm)), b = desc_sw128(smem_u32(sm + 16384));
        long long t0 = clock64();
        for (int i = 0; i < iters; i += BATCH) {
            // probe a barrier whose awaited phase is already complete (parity 1 on a fresh barrier)
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 1;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&ready)) : "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < BATCH; ++k) mma_ss(tb, a + 2 * (k & 3), b + 2 * (k & 3), idesc, 1);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
            }
            __syncwarp();
        }
        if (elect_one()) {
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
            cycles[blockIdx.x] = clock64() - t0;
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

template <int BATCH>
void run_batch(int grid) {
    const int iters = 24576;
    long long *d;
    cudaMalloc(&d, grid * sizeof(long long));
    const size_t smem = 32768 + 2048;
    cudaFuncSetAttribute(batch_kernel<BATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    batch_kernel<BATCH><<<grid, 128, smem>>>(iters, d);
    batch_kernel<BATCH><<<grid, 128, smem>>>(iters, d);
    cudaError_t err = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, d, sizeof(long long), cudaMemcpyDeviceToHost);
    printf("batches of %2d MMAs + commit/probe/fence/elect: %7.1f cycles/MMA  [%s]\n", BATCH, (double)h / iters,
           cudaGetErrorString(err));
    cudaFree(d);
}

template <int N, int MODE>
void run(const char *name, int grid) {
    const int iters = 65536;
    long long *d;
    cudaMalloc(&d, grid * sizeof(long long));
    const size_t smem = 16384 + N * 128 + 2048;
    cudaFuncSetAttribute(rate_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate_kernel<N, MODE><<<grid, 128, smem>>>(iters, d);
    cudaEventRecord(e0);
    rate_kernel<N, MODE><<<grid, 128, smem>>>(iters, d);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148];
    cudaMemcpy(h, d, (grid < 148 ? grid : 148) * sizeof(long long), cudaMemcpyDeviceToHost);
    const double flop = 2.0 * 128 * N * 16 * iters * grid;
    printf("%-44s grid %3d: %7.1f cycles/MMA (ideal %3d)  %8.1f TFLOP/s  [%s]\n", name, grid, (double)h[0] / iters,
           N / 2, flop / (ms * 1e-3) / 1e12, cudaGetErrorString(err));
    cudaFree(d);
}

int main() {
    run_batch<4>(148); run_batch<8>(148); run_batch<12>(148); run_batch<24>(148);
    for (int rnd : {0, 1}) {
    cudaMemcpyToSymbol(g_random_fill, &rnd, sizeof(int));
    printf("---- operands: %s\n", rnd ? "pseudo-random bf16" : "constant");
    for (int grid : {148}) {
        run<128, 0>("SS N=128 same accumulator", grid);
        run<128, 1>("SS N=128 two alternating accumulators", grid);
        run<128, 3>("SS N=128 four alternating accumulators", grid);
        run<128, 2>("TS N=128 same accumulator", grid);
        run<256, 0>("SS N=256 same accumulator", grid);
        run<256, 1>("SS N=256 two alternating accumulators", grid);
        run<64, 0>("SS N=64 same accumulator", grid);
        run<64, 1>("SS N=64 two alternating accumulators", grid);
    }
    }
    return 0;
}
