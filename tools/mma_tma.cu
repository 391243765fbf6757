// Microbenchmark: does a concurrent TMA stream into shared memory slow tcgen05.mma down?
// warp 1 issues MMAs (M128 N128 K16, bf16) whose B operand cycles over 12 ring slots; warp 0
// optionally streams 16 KB TMA boxes into the same slots as fast as they complete.  No data
// dependence between the two (operands are garbage) -- only the shared-memory traffic matters.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
    return (uint64_t)((a >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
constexpr int STAGES = 8, CH = 16384;  // + 8 more chunks after the ring for the resident strip (a_mode 3)

// tma_mode: 0 none, 1 stream at full rate, 2 stream paced to ~one 16 KB box per 384 cycles
// a_mode: 0 A from smem (SS), 1 A from TMEM (TS)
__global__ void __launch_bounds__(96, 1) k(const __grid_constant__ CUtensorMap tm, int iters, int tma_mode, int a_mode,
                                           int rows_total, long long *cycles, volatile int *stop_flag_unused) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[STAGES], done_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int s_stop;
    uint8_t *sm = (uint8_t *)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (STAGES + 1 + 4) * CH / 4; i += 96) ((uint32_t *)sm)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_stop = 0;
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | (8u << 24);
    if (warp == 0 && threadIdx.x == 0 && tma_mode) {
        int row = (blockIdx.x * 977) % (rows_total / 128) * 128;
        long long next = clock64();
        for (int i = 0; !s_stop; ++i) {
            const int s = i % STAGES;
            if (i >= STAGES) mbar_wait(smem_u32(&full[s]), ((i / STAGES) - 1) & 1);
            if (tma_mode == 2) { while (clock64() < next) {} next += 384; }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(CH) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(smem_u32(sm + CH + s * CH)), "l"(&tm), "r"(smem_u32(&full[s])), "r"((i & 3) * 64), "r"(row) : "memory");
            row += 128 * 148;
            if (row + 128 > rows_total) row -= rows_total / 256 * 128;
            if (row < 0) row = 0;
        }
        // drain: wait for everything in flight before the CTA exits
        for (int s = 0; s < STAGES; ++s) { /* best effort: spin a little */ }
        long long t = clock64(); while (clock64() - t < 20000) {}
    }
    if (warp == 1) {
        if (elect_one()) {
            const uint64_t a = desc_sw128(smem_u32(sm));
            long long t0 = clock64();
            // a_mode: 0 SS fixed A slices, 1 TS fixed A slices, 2 TS with the GEMM's A pattern (hi/lo, 4 K-chunks),
            //         3 SS with A cycling over 8 smem chunks (hi/lo x 4 K-chunks resident strip)
            const uint32_t ring = smem_u32(sm + CH);
            for (int i = 0; i < iters; i += 8) {
                const int st = (i >> 3) & 7;           // 8 B slots, power of two: no modulo on the issue path
                const int kc = (i >> 3) & 3;
                const uint64_t b = desc_sw128(ring + st * CH);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                    for (int hl = 0; hl < 2; ++hl) {
                        if (a_mode == 0)
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                         ::"r"(tb), "l"(a + 2 * kk), "l"(b + 2 * kk), "r"(idesc), "r"(1) : "memory");
                        else if (a_mode == 3)
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                         ::"r"(tb), "l"(desc_sw128(ring + (8 + ((kc + 4 * hl) & 3)) * CH) + 2 * kk), "l"(b + 2 * kk), "r"(idesc), "r"(1) : "memory");
                        else if (a_mode == 1)
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                         ::"r"(tb), "r"(tb + 256 + 8 * kk), "l"(b + 2 * kk), "r"(idesc), "r"(1) : "memory");
                        else
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                         ::"r"(tb), "r"(tb + 256 + 128 * hl + 32 * kc + 8 * kk), "l"(b + 2 * kk), "r"(idesc), "r"(1) : "memory");
                    }
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
            mbar_wait(smem_u32(&done_bar), 0);
            cycles[blockIdx.x] = clock64() - t0;
            s_stop = 1;
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeFn enc = (EncodeFn)fn;
    const int rows_total = 2 * 16 * 2176;
    void *buf;
    cudaMalloc(&buf, (size_t)rows_total * 512);
    cudaMemset(buf, 0x3c, (size_t)rows_total * 512);
    CUtensorMap tm;
    cuuint64_t dims[2] = {256, (cuuint64_t)rows_total};
    cuuint64_t strides[1] = {512};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t es[2] = {1, 1};
    enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    long long *d;
    cudaMalloc(&d, 148 * sizeof(long long));
    const size_t smem = (size_t)(STAGES + 1 + 4) * CH + 2048;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int iters = 16384;
    for (int a_mode : {0, 3, 1, 2})
        for (int tma_mode : {0, 2}) {
            k<<<148, 96, smem>>>(tm, iters, tma_mode, a_mode, rows_total, d, nullptr);
            cudaError_t err = cudaDeviceSynchronize();
            long long h = 0;
            cudaMemcpy(&h, d, sizeof(long long), cudaMemcpyDeviceToHost);
            const char *an[4] = {"smem, 4 fixed slices", "TMEM, 4 fixed slices", "TMEM, GEMM pattern (256 cols)", "smem, cycling 4 chunks"};
            printf("A from %-30s TMA %-34s: %6.1f cycles/MMA  [%s]\n", an[a_mode],
                   tma_mode == 0 ? "off" : tma_mode == 1 ? "full rate (~87 B/clk)" : "paced 16 KB / 384 clk (43 B/clk)",
                   (double)h / iters, cudaGetErrorString(err));
        }
    return 0;
}
