// Shared device/host helpers for libpem_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include <atomic>

#include "pem_b200.h"

#define PEM_NUM_SMS 148  // B200: 2 dies x 74 SMs

#define PEM_REQUIRE_PTR(p) \
    do {                   \
        if ((p) == nullptr) return PEM_E_NULL; \
    } while (0)
#define PEM_REQUIRE(cond, code) \
    do {                        \
        if (!(cond)) return (code); \
    } while (0)

// Launch-status check: never synchronises; a sticky earlier error is also reported.
static inline int pem_launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? PEM_OK : (int)e;
}

// Opt a kernel into `bytes` of dynamic shared memory.  The attribute is per DEVICE, so the sizes already
// granted are cached per call site AND per device (relaxed atomics: host threads may race to raise it, the
// worst case is one redundant cudaFuncSetAttribute).  Steady-state calls -- including calls made while the
// stream is being captured into a CUDA graph -- issue no runtime API call beyond cudaGetDevice.
#define PEM_MAX_DEVICES 64
struct PemSmemCache {
    std::atomic<size_t> granted[PEM_MAX_DEVICES];
    explicit PemSmemCache(size_t initial = 0) {
        for (auto &g : granted) g.store(initial, std::memory_order_relaxed);
    }
};
static inline int pem_current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev;
}
template <typename Kern>
static inline int pem_ensure_dyn_smem(Kern kern, size_t bytes, PemSmemCache *cache) {
    const int dev = pem_current_device();
    const bool tracked = dev >= 0 && dev < PEM_MAX_DEVICES;
    if (tracked && bytes <= cache->granted[dev].load(std::memory_order_relaxed)) return PEM_OK;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    if (tracked) cache->granted[dev].store(bytes, std::memory_order_relaxed);
    return PEM_OK;
}
// Number of SMs of the current device (cached per device).
static inline int pem_num_sms() {
    static std::atomic<int> sms[PEM_MAX_DEVICES];
    const int dev = pem_current_device();
    const bool tracked = dev >= 0 && dev < PEM_MAX_DEVICES;
    int n = tracked ? sms[dev].load(std::memory_order_relaxed) : 0;
    if (n <= 0) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = PEM_NUM_SMS;
        if (tracked) sms[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

static inline cudaStream_t pem_cu_stream(pem_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch (PDL).  A kernel launched through pem_launch_pdl may be scheduled while
// its predecessor on the stream is still running: its CTAs do their local prologue (barrier
// initialisation, TMEM allocation, index arithmetic) and park in pdl_enter() until the predecessor has
// completed and flushed, so launch latency and prologue hide under the predecessor's tail.  EVERY kernel
// launched this way calls pem::pdl_enter() before its first global-memory access; a kernel that calls it
// under a plain launch passes straight through.  Works across stream capture (the graph keeps the
// programmatic edge).  In a -DPEM_DEV_KNOBS build pem_dev_set_pdl(0) turns such launches back into plain ones.
#ifdef PEM_DEV_KNOBS
extern int g_pem_pdl;  // pem_api.cu
#else
static constexpr int g_pem_pdl = 1;
#endif
template <typename... KArgs, typename... Args>
static inline void pem_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                  Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pem_pdl ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);  // status: pem_launch_status()
}

namespace pem {

constexpr unsigned kFull = 0xffffffffu;

// see pem_launch_pdl: wait for the predecessor grid (complete + visible), then let the successor be scheduled
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// Unfused squared distance in the reference CPU evaluation order
// ((dx*dx)+(dy*dy))+(dz*dz): nvcc must not contract these into FMAs, otherwise
// sampled indices / neighbour lists stop being bit-exact.
__device__ __forceinline__ float sqdist_unfused(float ax, float ay, float az, float bx, float by,
                                                float bz) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Block-wide sum for blocks of up to 1024 threads; result valid in every thread.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T *smem /* >= 32 entries */) {
    v = warp_sum(v);
    const int w = warp_id(), l = lane_id(), nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) smem[w] = v;
    __syncthreads();
    T r = (l < nw) ? smem[l] : T(0);
    r = warp_sum(r);
    return r;
}

// N block-wide sums behind ONE pair of barriers (fixed order: deterministic); results in every thread.
// smem: N rows of 33 entries.
template <int N, typename T>
__device__ __forceinline__ void block_sum_n(T (&v)[N], T (*smem)[33]) {
    const int w = warp_id(), l = lane_id(), nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();
    if (l == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) smem[k][w] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; ++k) {
        T r = (l < nw) ? smem[k][l] : T(0);
        v[k] = warp_sum(r);
    }
}

}  // namespace pem
