// Multi-GPU pose-record gather without a collective call ("fused variant"): the kernel that packs
// the 64-byte per-instance records stores them straight into EVERY rank's record table over NVLink (peer
// pointers obtained once through CUDA IPC) and raises a per-rank flag; a second tiny kernel on each rank waits
// for the flags of the step and hands the consumer a stable copy of the table.  Both kernels are ordinary
// stream work, so they live INSIDE the step's CUDA graph: a step at N GPUs is the same single graph launch as
// at 1 GPU (the NCCL all-gather had to follow the graph replay as a separate launch).  The reference has no
// counterpart (its inference path is single-device).
//
// Per-rank table (same layout on every rank, `pem_peer_table_bytes`):
//   header   256 B   [0] step counter of THIS rank   [1] error flag (wait timed out)
//   flags    2 x kMaxPeers x 4 B   flags[parity][r] = last step of parity `parity` whose records rank r delivered
//   records  2 x (world * n_per_rank) x 64 B          double buffered by step parity
// A rank can run at most one step ahead of a peer (its wait for step s+1 needs the peer's flag of s+1, which the
// peer raises only after consuming step s on its stream), so two buffers suffice.
#include "pem_common.cuh"

namespace pem {

constexpr int kMaxPeers = 16;
constexpr size_t kPeerHeader = 256, kPeerFlags = 2 * kMaxPeers * sizeof(unsigned);

struct PeerPtrs { unsigned char *p[kMaxPeers]; };

__host__ __device__ inline size_t peer_records_offset(int parity, int world, int n_per_rank) {
    return kPeerHeader + kPeerFlags + (size_t)parity * world * n_per_rank * 64;
}

__global__ void __launch_bounds__(256)
gather_poses_kernel(const float *__restrict__ R, const float *__restrict__ t, const float *__restrict__ score,
                    const int32_t *__restrict__ pid, const float *__restrict__ inst, int B, int rank, int world,
                    PeerPtrs peers, float *__restrict__ rec_local) {
    unsigned *hdr = reinterpret_cast<unsigned *>(peers.p[rank]);
    const unsigned step = hdr[0] + 1u;  // every thread reads before thread 0 advances it (barrier below)
    const int parity = step & 1u;
    const size_t off = peer_records_offset(parity, world, B) + (size_t)rank * B * 64;
    for (int i = threadIdx.x; i < B * 16; i += blockDim.x) {
        const int b = i >> 4, k = i & 15;
        float v = 0.f;
        if (k < 9) v = R[b * 9 + k];
        else if (k < 12) v = t[b * 3 + k - 9];
        else if (k == 12) v = score[b];
        else if (k == 13) v = pid ? (float)pid[b] : 0.f;
        else if (k == 14) v = inst ? inst[b] : (float)b;
        if (rec_local) rec_local[i] = v;
        for (int p = 0; p < world; ++p)  // coalesced 64-byte rows into every rank's table (own included)
            reinterpret_cast<float *>(peers.p[p] + off)[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
        unsigned *flag = reinterpret_cast<unsigned *>(peers.p[threadIdx.x] + kPeerHeader) + parity * kMaxPeers + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(step) : "memory");
    }
    if (threadIdx.x == 0) hdr[0] = step;
}

__global__ void __launch_bounds__(256)
gather_poses_wait_kernel(unsigned char *__restrict__ own, int world, int n_per_rank, float *__restrict__ out,
                         unsigned long long timeout_ns) {
    unsigned *hdr = reinterpret_cast<unsigned *>(own);
    const unsigned step = hdr[0];  // advanced by gather_poses_kernel earlier on this stream
    const int parity = step & 1u;
    if (threadIdx.x < world) {
        const unsigned *flag = reinterpret_cast<const unsigned *>(own + kPeerHeader) + parity * kMaxPeers + threadIdx.x;
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            unsigned v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            if ((int)(v - step) >= 0) break;
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) {  // a peer never delivered: flag the error instead of hanging the GPU
                hdr[1] = 1u;
                break;
            }
            __nanosleep(200);
        }
    }
    __syncthreads();
    if (out) {
        const float4 *src = reinterpret_cast<const float4 *>(own + peer_records_offset(parity, world, n_per_rank));
        float4 *dst = reinterpret_cast<float4 *>(out);
        for (int i = threadIdx.x; i < world * n_per_rank * 4; i += blockDim.x) dst[i] = src[i];
    }
}

}  // namespace pem

using namespace pem;

extern "C" size_t pem_peer_table_bytes(int world, int n_per_rank) {
    if (world <= 0 || world > kMaxPeers || n_per_rank <= 0) return 0;
    return peer_records_offset(2, world, n_per_rank);
}

// One-time setup (these three DO call the allocator / IPC API; the per-step entry points below do not).
extern "C" int pem_peer_alloc(size_t bytes, void **ptr, unsigned char *ipc_handle /* 64 bytes */) {
    PEM_REQUIRE_PTR(ptr);
    PEM_REQUIRE_PTR(ipc_handle);
    PEM_REQUIRE(bytes > 0, PEM_E_SHAPE);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(*ptr, 0, bytes);
    if (e != cudaSuccess) return (int)e;
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, *ptr);
    if (e != cudaSuccess) return (int)e;
    memcpy(ipc_handle, &h, 64);
    return PEM_OK;
}

extern "C" int pem_peer_open(const unsigned char *ipc_handle, void **ptr) {
    PEM_REQUIRE_PTR(ptr);
    PEM_REQUIRE_PTR(ipc_handle);
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, 64);
    cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
    return e == cudaSuccess ? PEM_OK : (int)e;
}

extern "C" int pem_peer_release(void *ptr, int opened) {
    if (ptr == nullptr) return PEM_OK;
    cudaError_t e = opened ? cudaIpcCloseMemHandle(ptr) : cudaFree(ptr);
    return e == cudaSuccess ? PEM_OK : (int)e;
}

// Pack this rank's records (like pem_pack_records, also into `records_local` when given) and deliver them to
// the tables of all `world` ranks.  peer_tables: HOST array of `world` device pointers (peer_tables[rank] =
// this rank's own table), each of pem_peer_table_bytes(world, n_per_rank) bytes.
extern "C" int pem_gather_poses(const float *R, const float *t, const float *score, const int32_t *proposal_id,
                                const float *instance_id, int n_per_rank, int rank, int world,
                                void *const *peer_tables, float *records_local, pem_stream_t stream) {
    PEM_REQUIRE_PTR(R);
    PEM_REQUIRE_PTR(t);
    PEM_REQUIRE_PTR(score);
    PEM_REQUIRE_PTR(peer_tables);
    PEM_REQUIRE(n_per_rank > 0 && world > 0 && world <= kMaxPeers && rank >= 0 && rank < world, PEM_E_SHAPE);
    PeerPtrs pp{};
    for (int p = 0; p < world; ++p) {
        PEM_REQUIRE_PTR(peer_tables[p]);
        pp.p[p] = (unsigned char *)peer_tables[p];
    }
    gather_poses_kernel<<<1, 256, 0, pem_cu_stream(stream)>>>(R, t, score, proposal_id, instance_id, n_per_rank, rank,
                                                             world, pp, records_local);
    return pem_launch_status();
}

// Wait until every rank's records of the current step have arrived in `own_table`, then copy the step's
// (world * n_per_rank, 16) table to `out` (may be NULL: wait only).  A peer that does not deliver within
// timeout_ms sets the table's error flag (header word 1) instead of hanging the device.
extern "C" int pem_gather_poses_wait(void *own_table, int world, int n_per_rank, float *out, int timeout_ms,
                                     pem_stream_t stream) {
    PEM_REQUIRE_PTR(own_table);
    PEM_REQUIRE(n_per_rank > 0 && world > 0 && world <= kMaxPeers && timeout_ms > 0, PEM_E_SHAPE);
    gather_poses_wait_kernel<<<1, 256, 0, pem_cu_stream(stream)>>>((unsigned char *)own_table, world, n_per_rank, out,
                                                                  (unsigned long long)timeout_ms * 1000000ull);
    return pem_launch_status();
}

// fp16 -> fp32 widening of an uploaded feature tensor (the host-facing path may upload half-precision
// features: half the PCIe bytes); n must be a multiple of 8, pointers 16-byte aligned.
namespace pem {
__global__ void __launch_bounds__(256)
half_to_float_kernel(const uint4 *__restrict__ src, float4 *__restrict__ dst, size_t n8) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(src + i);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
        float f[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            asm("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %2; cvt.f32.f16 %0, lo; cvt.f32.f16 %1, hi;}"
                : "=f"(f[2 * k]), "=f"(f[2 * k + 1]) : "r"(w[k]));
        }
        dst[2 * i] = make_float4(f[0], f[1], f[2], f[3]);
        dst[2 * i + 1] = make_float4(f[4], f[5], f[6], f[7]);
    }
}
}  // namespace pem

extern "C" int pem_half_to_float(const void *src, float *dst, size_t n, pem_stream_t stream) {
    PEM_REQUIRE_PTR(src);
    PEM_REQUIRE_PTR(dst);
    PEM_REQUIRE(n > 0 && n % 8 == 0, PEM_E_SHAPE);
    PEM_REQUIRE((uintptr_t)src % 16 == 0 && (uintptr_t)dst % 16 == 0, PEM_E_UNSUPPORTED);
    const size_t n8 = n / 8;
    const int grid = (int)min((size_t)pem_num_sms() * 8, (n8 + 255) / 256);
    half_to_float_kernel<<<grid, 256, 0, pem_cu_stream(stream)>>>((const uint4 *)src, (float4 *)dst, n8);
    return pem_launch_status();
}
