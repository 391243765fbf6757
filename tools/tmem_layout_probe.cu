// Which (lane, column) does register i of thread t receive from tcgen05.ld.16x256b.x2 ?  TMEM is filled through
// tcgen05.st.32x32b (thread = lane) with lane * 1000 + column, then read back in the 16x256b shape at lane offsets 0 and 16.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_layout_probe tools/tmem_layout_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(uint32_t *out) {
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tbase;
    uint32_t v[16];
    for (int c = 0; c < 16; ++c) v[c] = (uint32_t)(threadIdx.x * 1000 + c);
    const uint32_t ta = base + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(ta),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                 "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int lh = 0; lh < 2; ++lh) {
        uint32_t r[8];
        const uint32_t tb = base + ((uint32_t)(warp * 32 + lh * 16) << 16);
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(tb) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) out[((warp * 2 + lh) * 32 + lane) * 8 + i] = r[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(32u) : "memory");
}
int main() {
    uint32_t *d, h[4 * 2 * 32 * 8];
    cudaMalloc(&d, sizeof(h));
    cudaMemset(d, 0xff, sizeof(h));
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int w = 0; w < 2; ++w)
        for (int lh = 0; lh < 2; ++lh)
            for (int t = 0; t < 32; t += (t < 8 ? 1 : 7)) {
                printf("warp %d lane-offset %2d thread %2d:", w, 16 * lh, t);
                for (int i = 0; i < 8; ++i) { uint32_t x = h[((w * 2 + lh) * 32 + t) * 8 + i]; printf("  (l%u,c%u)", x / 1000, x % 1000); }
                printf("\n");
            }
    return 0;
}
