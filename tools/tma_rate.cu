// Microbenchmark: TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B boxes of 64 bf16 x ROWS) ingest rate per
// SM from an L2-resident tensor through a 12-slot ring, no consumer work (slots are recycled as soon
// as they land).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_rate tools/tma_rate.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

template <int ROWS, int STAGES>
__global__ void __launch_bounds__(64, 1) tma_kernel(const __grid_constant__ CUtensorMap tm, int n_loads, int rows_total,
                                                    long long *cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[STAGES];
    uint8_t *sm = (uint8_t *)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    constexpr int BYTES = ROWS * 128;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        int row = (blockIdx.x * 977) % (rows_total / ROWS) * ROWS;
        // prime the ring, then: wait for a slot to land, immediately reload it
        for (int i = 0; i < n_loads + STAGES; ++i) {
            const int s = i % STAGES;
            if (i >= STAGES) mbar_wait(smem_u32(&full[s]), ((i / STAGES) - 1) & 1);
            if (i < n_loads) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(BYTES) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(smem_u32(sm + s * BYTES)), "l"(&tm), "r"(smem_u32(&full[s])), "r"((i & 3) * 64), "r"(row) : "memory");
                row += ROWS * 148;
                if (row + ROWS > rows_total) row -= rows_total / ROWS / 2 * ROWS;
                if (row < 0) row = 0;
            }
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int ROWS, int STAGES>
void run(EncodeFn enc, void *buf, int rows_total, int grid) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {256, (cuuint64_t)rows_total};
    cuuint64_t strides[1] = {512};
    cuuint32_t box[2] = {64, ROWS};
    cuuint32_t es[2] = {1, 1};
    enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int n_loads = 4096;
    long long *d;
    cudaMalloc(&d, grid * sizeof(long long));
    const size_t smem = (size_t)STAGES * ROWS * 128 + 2048;
    cudaFuncSetAttribute(tma_kernel<ROWS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    tma_kernel<ROWS, STAGES><<<grid, 64, smem>>>(tm, n_loads, rows_total, d);
    cudaEventRecord(e0);
    tma_kernel<ROWS, STAGES><<<grid, 64, smem>>>(tm, n_loads, rows_total, d);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long h = 0;
    cudaMemcpy(&h, d, sizeof(long long), cudaMemcpyDeviceToHost);
    const double bytes = (double)n_loads * ROWS * 128;
    printf("box %3d rows (%5d B) x %2d slots, grid %3d: %6.1f B/clk/SM, %7.2f TB/s chip  [%s]\n", ROWS, ROWS * 128, STAGES, grid,
           bytes / (double)h, bytes * grid / (ms * 1e-3) / 1e12, cudaGetErrorString(err));
    cudaFree(d);
}

int main() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeFn enc = (EncodeFn)fn;
    const int rows_total = 2 * 16 * 2176;  // 35.6 MB like one prepared operand (L2 resident)
    void *buf;
    cudaMalloc(&buf, (size_t)rows_total * 512);
    cudaMemset(buf, 0x11, (size_t)rows_total * 512);
    for (int grid : {1, 148}) {
        run<128, 12>(enc, buf, rows_total, grid);
        run<128, 6>(enc, buf, rows_total, grid);
        run<128, 3>(enc, buf, rows_total, grid);
        run<64, 12>(enc, buf, rows_total, grid);
        run<64, 24>(enc, buf, rows_total, grid);
        run<256, 6>(enc, buf, rows_total, grid);
    }
    return 0;
}
