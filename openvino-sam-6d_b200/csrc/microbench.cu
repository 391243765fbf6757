// FP32 issue-peak probe ("FP32 peak: measure an FMA-saturating microbench on the box").  The
// pointnet2 / scoring kernels are bound by the FP32 pipe, not by HBM, so their roofline denominator is what the
// 148 SMs sustain in (a) fused multiply-adds (2 flop per lane-instruction) and (b) the UNFUSED mul / add / sub
// / min / compare stream that bit-exactness with the reference's CPU loops mandates (1 op per lane-instruction).
// bench.py times this kernel with CUDA events and reports both under extra.fp32_peak.
#include "pem_common.cuh"

namespace pem {

template <bool FUSED>
__global__ void __launch_bounds__(512)
fp32_peak_kernel(int iters, float seed, float *__restrict__ sink) {
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + (float)(threadIdx.x + k);
    const float m = 0.999f + seed, c = 1e-3f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (FUSED) a[k] = fmaf(a[k], m, c);
                else a[k] = (r & 1) ? __fadd_rn(a[k], c) : __fmul_rn(a[k], m);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 12345.678f) sink[0] = s;  // never true; keeps the chains alive
}

// packed pairs: fma.rn.f32x2 (FFMA2), two fused multiply-adds per lane-instruction
__global__ void __launch_bounds__(512)
fp32x2_peak_kernel(int iters, float seed, float *__restrict__ sink) {
    unsigned long long a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float lo = seed + (float)(threadIdx.x + k), hi = lo + 0.5f;
        asm("mov.b64 %0, {%1, %2};" : "=l"(a[k]) : "f"(lo), "f"(hi));
    }
    unsigned long long m, c;
    {
        const float mf = 0.999f + seed, cf = 1e-3f;
        asm("mov.b64 %0, {%1, %1};" : "=l"(m) : "f"(mf));
        asm("mov.b64 %0, {%1, %1};" : "=l"(c) : "f"(cf));
    }
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int k = 0; k < 8; ++k) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[k]) : "l"(m), "l"(c));
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s ^= a[k];
    if (s == 0x123456789abcdefull) sink[0] = 1.f;
}

}  // namespace pem

// mode 0: FFMA chains, mode 1: alternating FMUL / FADD, mode 2: FFMA2 chains (4 flop per lane-instruction).  Lane-instructions executed = grid * 512 * iters * 64.
extern "C" int pem_bench_fp32_peak(int mode, int ctas, int iters, float *sink, pem_stream_t stream) {
    PEM_REQUIRE_PTR(sink);
    PEM_REQUIRE(ctas > 0 && iters > 0 && mode >= 0 && mode <= 2, PEM_E_SHAPE);
    if (mode == 0) pem::fp32_peak_kernel<true><<<ctas, 512, 0, pem_cu_stream(stream)>>>(iters, 0.f, sink);
    else if (mode == 2) pem::fp32x2_peak_kernel<<<ctas, 512, 0, pem_cu_stream(stream)>>>(iters, 0.f, sink);
    else pem::fp32_peak_kernel<false><<<ctas, 512, 0, pem_cu_stream(stream)>>>(iters, 0.f, sink);
    return pem_launch_status();
}
