// pointnet2 ops of the PEM hot path as sm_100a kernels: furthest point sampling,
// gather, ball query, grouping and the fused QueryAndGroup.
//
// Reference semantics (bit-exact targets) -- paths relative to
// /root/reference/SAM-6D/Pose_Estimation_Model/model/pointnet2/_ext_src/src/ :
//   FPS         sampling.cpp:76-118 (CPU loop; CUDA twin sampling_gpu.cu:75-178)
//   gather      sampling.cpp:23-44  (CUDA twin sampling_gpu.cu:13-25)
//   ball query  ball_query.cpp:16-62 (CUDA twin ball_query_gpu.cu:14-49)
//   grouping    group_points.cpp:20-45 (CUDA twin group_points_gpu.cu:13-33)
//   QueryAndGroup  ../pointnet2_utils.py:326-403
//
// Design notes (B200):
//  * FPS is a chain of M-1 dependent argmax steps, so it is latency/issue bound,
//    not HBM bound.  The cloud is read from HBM exactly once into shared memory
//    (SoA, conflict-free) and -- up to 8192 points -- into registers together with
//    the running min-distance; each step is one pass over register data, a
//    redux.sync warp argmax, ONE __syncthreads (double-buffered per-warp keys) and
//    a redundant per-warp final reduce.  Ties resolve to the lowest index by
//    reducing the pair (distance bits, ~index) as an unsigned maximum.
//  * Ball query is order dependent (first nsample hits in index order), so each
//    thread owns one centre and scans the points in order from a shared-memory
//    tile (one LDS.128 broadcast per point per warp); neighbour rows are staged in
//    shared memory and written coalesced.  The fused QueryAndGroup consumes the
//    staged rows directly and writes the (B,3+C,M,ns) tensor once.
#include <cooperative_groups.h>

#include "pem_common.cuh"

namespace cg = cooperative_groups;

namespace pem {

// ------------------------------------------------------------------------------------
// FPS
// ------------------------------------------------------------------------------------
__device__ __forceinline__ bool near_origin(float x, float y, float z) {
    // sampling.cpp:101-103: float mag compared with the double literal 1e-3
    const float mag = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    return (double)mag <= 1e-3;
}

// argmax of (value, lowest index) over the CTA; returns the winning index (0 if no
// thread has a valid candidate).  hi = float bits of a non-negative value (0 when
// invalid), lo = ~index (0 when invalid).
template <int THREADS>
__device__ __forceinline__ int cta_argmax(unsigned hi, unsigned lo, unsigned long long *s_key) {
    constexpr int NW = THREADS / 32;
    const unsigned whi = __reduce_max_sync(kFull, hi);
    const unsigned wlo = __reduce_max_sync(kFull, hi == whi ? lo : 0u);
    if (NW == 1) return wlo ? (int)~wlo : 0;
    if (lane_id() == 0) s_key[warp_id()] = ((unsigned long long)whi << 32) | wlo;
    __syncthreads();
    const unsigned long long key = lane_id() < NW ? s_key[lane_id()] : 0ull;
    const unsigned h2 = (unsigned)(key >> 32), l2 = (unsigned)key;
    const unsigned ghi = __reduce_max_sync(kFull, h2);
    const unsigned glo = __reduce_max_sync(kFull, h2 == ghi ? l2 : 0u);
    return glo ? (int)~glo : 0;
}

// One CTA per cloud, N <= THREADS*PPT.  IN_REGS: coordinates live in registers;
// otherwise they are re-read from the SoA shared-memory copy every step.
template <int THREADS, int PPT, bool IN_REGS>
__global__ void __launch_bounds__(THREADS, 1)
fps_cta_kernel(const float *__restrict__ xyz, int N, int M, int skip, int32_t *__restrict__ idx,
               float *__restrict__ new_xyz /* (B,M,3) or null: the sampled points themselves */) {
    extern __shared__ float s_soa[];  // x[N] | y[N] | z[N]
    __shared__ unsigned long long s_key[2][32];
    float *sx = s_soa, *sy = s_soa + N, *sz = s_soa + 2 * N;
    const int tid = threadIdx.x;
    const float *src = xyz + (size_t)blockIdx.x * N * 3;
    int32_t *out = idx + (size_t)blockIdx.x * M;
    float *gout = new_xyz ? new_xyz + (size_t)blockIdx.x * M * 3 : nullptr;

    for (int e = tid; e < 3 * N; e += THREADS) {
        const int k = e / 3, c = e - 3 * k;
        s_soa[c * N + k] = src[e];
    }
    __syncthreads();

    float px[IN_REGS ? PPT : 1], py[IN_REGS ? PPT : 1], pz[IN_REGS ? PPT : 1];
    float run[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int k = i * THREADS + tid;
        bool ok = k < N;
        float x = 0.f, y = 0.f, z = 0.f;
        if (ok) {
            x = sx[k]; y = sy[k]; z = sz[k];
            if (skip && near_origin(x, y, z)) ok = false;
        }
        if (IN_REGS) { px[i] = x; py[i] = y; pz[i] = z; }
        // -inf marks "never selectable": fminf(d, -inf) stays -inf and -inf > best is false
        run[i] = ok ? FLT_MAX : -INFINITY;
    }

    int cur = 0;
    if (tid == 0) out[0] = 0;
    for (int j = 1; j < M; ++j) {
        const float cx = sx[cur], cy = sy[cur], cz = sz[cur];
        if (gout && tid == 0) {  // the gather of the sampled point rides along (its own kernel cost 6 us)
            gout[3 * (j - 1)] = cx; gout[3 * (j - 1) + 1] = cy; gout[3 * (j - 1) + 2] = cz;
        }
        float best = -FLT_MAX;
        int arg = 0;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const int k = i * THREADS + tid;
            float x, y, z;
            if (IN_REGS) { x = px[i]; y = py[i]; z = pz[i]; }
            else if (k < N) { x = sx[k]; y = sy[k]; z = sz[k]; }
            else { x = y = z = 0.f; }
            const float d = sqdist_unfused(x, y, z, cx, cy, cz);
            const float m = fminf(d, run[i]);
            run[i] = m;
            if (m > best) { best = m; arg = k; }  // ascending k + strict > : lowest index wins
        }
        const bool valid = best > -FLT_MAX;
        cur = cta_argmax<THREADS>(valid ? __float_as_uint(best) : 0u, valid ? ~(unsigned)arg : 0u,
                                  s_key[j & 1]);
        if (tid == 0) out[j] = cur;
    }
    if (gout && tid == 0) {
        gout[3 * (M - 1)] = sx[cur]; gout[3 * (M - 1) + 1] = sy[cur]; gout[3 * (M - 1) + 2] = sz[cur];
    }
}

// Thread-block-cluster variant for clouds too large for one SM's registers: the cloud is
// split into contiguous slabs, one per CTA of the cluster (coordinates + running distance in
// registers, or coordinates in that CTA's shared memory).  Every step each CTA reduces its slab
// locally, then writes a 20-byte record (distance bits, ~global index, x, y, z of its local
// winner) into every peer's shared memory through DSMEM; one cluster barrier later all CTAs
// hold all records, pick the same global winner and already have its coordinates.
struct FpsRecord { unsigned hi, lo; float x, y, z; };

template <int THREADS, int PPT, bool IN_REGS>
__global__ void __launch_bounds__(THREADS, 1)
fps_cluster_kernel(const float *__restrict__ xyz, int N, int M, int skip, int32_t *__restrict__ idx) {
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int cloud = blockIdx.x / C;
    const int per = (N + C - 1) / C;
    const int lo_pt = rank * per;
    const int n_loc = max(0, min(per, N - lo_pt));
    extern __shared__ float s_soa[];  // x[per] | y[per] | z[per] of this CTA's slab
    __shared__ unsigned long long s_key[2][32];
    __shared__ FpsRecord s_rec[2][16];
    float *sx = s_soa, *sy = s_soa + per, *sz = s_soa + 2 * per;
    const int tid = threadIdx.x;
    const float *src = xyz + ((size_t)cloud * N + lo_pt) * 3;
    int32_t *out = idx + (size_t)cloud * M;

    for (int e = tid; e < 3 * n_loc; e += THREADS) {
        const int k = e / 3, c = e - 3 * k;
        s_soa[c * per + k] = src[e];
    }
    __syncthreads();
    float px[IN_REGS ? PPT : 1], py[IN_REGS ? PPT : 1], pz[IN_REGS ? PPT : 1];
    float run[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int k = i * THREADS + tid;
        bool ok = k < n_loc;
        float x = 0.f, y = 0.f, z = 0.f;
        if (ok) {
            x = sx[k]; y = sy[k]; z = sz[k];
            if (skip && near_origin(x, y, z)) ok = false;
        }
        if (IN_REGS) { px[i] = x; py[i] = y; pz[i] = z; }
        run[i] = ok ? FLT_MAX : -INFINITY;
    }
    // the first sample is point 0 of the cloud (owned by rank 0): every CTA reads it from global
    const float *p0 = xyz + (size_t)cloud * N * 3;
    float cx = __ldg(p0), cy = __ldg(p0 + 1), cz = __ldg(p0 + 2);
    if (rank == 0 && tid == 0) out[0] = 0;
    cluster.sync();  // peers' shared memory is live before the first remote store
    for (int j = 1; j < M; ++j) {
        float best = -FLT_MAX;
        int arg = 0;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const int k = i * THREADS + tid;
            float x, y, z;
            if (IN_REGS) { x = px[i]; y = py[i]; z = pz[i]; }
            else if (k < n_loc) { x = sx[k]; y = sy[k]; z = sz[k]; }
            else { x = y = z = 0.f; }
            const float m = fminf(sqdist_unfused(x, y, z, cx, cy, cz), run[i]);
            run[i] = m;
            if (m > best) { best = m; arg = k; }
        }
        const bool valid = best > -FLT_MAX;
        // local winner (local index; 0 with hi = lo = 0 when the slab has no candidate)
        const unsigned hi = valid ? __float_as_uint(best) : 0u;
        const unsigned lo = valid ? ~(unsigned)(lo_pt + arg) : 0u;
        // CTA-level reduction of (hi, lo) -- same scheme as cta_argmax, but keep the pair
        constexpr int NW = THREADS / 32;
        const unsigned whi = __reduce_max_sync(kFull, hi);
        const unsigned wlo = __reduce_max_sync(kFull, hi == whi ? lo : 0u);
        const int par = j & 1;
        if (lane_id() == 0) s_key[par][warp_id()] = ((unsigned long long)whi << 32) | wlo;
        __syncthreads();
        // warp 0 computes the CTA key; lane r then delivers this CTA's record to CTA r
        if (warp_id() == 0) {
            const unsigned long long key = lane_id() < NW ? s_key[par][lane_id()] : 0ull;
            const unsigned h2 = (unsigned)(key >> 32), l2 = (unsigned)key;
            const unsigned chi = __reduce_max_sync(kFull, h2);
            const unsigned clo = __reduce_max_sync(kFull, h2 == chi ? l2 : 0u);
            if (lane_id() < C) {
                FpsRecord r;
                r.hi = chi; r.lo = clo;
                const int li = clo ? (int)(~clo) - lo_pt : 0;
                r.x = sx[li]; r.y = sy[li]; r.z = sz[li];
                FpsRecord *dst = cluster.map_shared_rank(&s_rec[par][rank], lane_id());
                *dst = r;
            }
        }
        cluster.sync();  // release/acquire: all records visible everywhere
        // global winner: max (hi, lo) over the C records; ties on distance -> lowest global index
        unsigned ghi = 0u, glo = 0u;
        float gx = 0.f, gy = 0.f, gz = 0.f;
        for (int r = 0; r < C; ++r) {
            const FpsRecord q = s_rec[par][r];
            if (q.hi > ghi || (q.hi == ghi && q.lo > glo)) { ghi = q.hi; glo = q.lo; gx = q.x; gy = q.y; gz = q.z; }
        }
        int cur = 0;
        if (glo != 0u) { cur = (int)~glo; cx = gx; cy = gy; cz = gz; }
        else { cx = __ldg(p0); cy = __ldg(p0 + 1); cz = __ldg(p0 + 2); }  // nothing selectable: index 0
        if (rank == 0 && tid == 0) out[j] = cur;
    }
    cluster.sync();  // no CTA exits while a peer may still write into its shared memory
}

// Fallback for clouds that do not fit one SM: coordinates from L2, running distance in
// a caller-provided workspace (B*N floats).
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
fps_stream_kernel(const float *__restrict__ xyz, int N, int M, int skip, float *__restrict__ run,
                  int32_t *__restrict__ idx) {
    __shared__ unsigned long long s_key[2][32];
    const int tid = threadIdx.x;
    const float *src = xyz + (size_t)blockIdx.x * N * 3;
    float *r = run + (size_t)blockIdx.x * N;
    int32_t *out = idx + (size_t)blockIdx.x * M;
    for (int k = tid; k < N; k += THREADS) {
        const bool bad = skip && near_origin(src[3 * k], src[3 * k + 1], src[3 * k + 2]);
        r[k] = bad ? -INFINITY : FLT_MAX;
    }
    int cur = 0;
    if (tid == 0) out[0] = 0;
    for (int j = 1; j < M; ++j) {
        const float cx = __ldg(src + 3 * cur), cy = __ldg(src + 3 * cur + 1),
                    cz = __ldg(src + 3 * cur + 2);
        float best = -FLT_MAX;
        int arg = 0;
        for (int k = tid; k < N; k += THREADS) {
            const float d = sqdist_unfused(__ldg(src + 3 * k), __ldg(src + 3 * k + 1),
                                           __ldg(src + 3 * k + 2), cx, cy, cz);
            const float m = fminf(d, r[k]);
            r[k] = m;
            if (m > best) { best = m; arg = k; }
        }
        const bool valid = best > -FLT_MAX;
        cur = cta_argmax<THREADS>(valid ? __float_as_uint(best) : 0u, valid ? ~(unsigned)arg : 0u,
                                  s_key[j & 1]);
        if (tid == 0) out[j] = cur;
    }
}

template <int THREADS, int PPT, bool IN_REGS>
static int launch_fps_cta(const float *xyz, int B, int N, int M, int skip, int32_t *idx,
                          cudaStream_t st, float *new_xyz = nullptr) {
    const size_t smem = (size_t)3 * N * sizeof(float);
    auto kern = fps_cta_kernel<THREADS, PPT, IN_REGS>;
    static PemSmemCache cached(40 * 1024);  // static shared memory counts against the 48 KB default too
    if (int rc = pem_ensure_dyn_smem(kern, smem, &cached)) return rc;
    kern<<<B, THREADS, smem, st>>>(xyz, N, M, skip, idx, new_xyz);
    return pem_launch_status();
}

template <int THREADS, int PPT, bool IN_REGS>
static int launch_fps_cluster(const float *xyz, int B, int N, int M, int skip, int32_t *idx, int C,
                              cudaStream_t st) {
    const int per = (N + C - 1) / C;
    const size_t smem = (size_t)3 * per * sizeof(float);
    auto kern = fps_cluster_kernel<THREADS, PPT, IN_REGS>;
    static PemSmemCache cached(40 * 1024);
    if (int rc = pem_ensure_dyn_smem(kern, smem, &cached)) return rc;
    static PemSmemCache nonportable(0);  // per device, like the shared-memory opt-in (1 = attribute set)
    if (C > 8) {
        const int dev = pem_current_device();
        const bool tracked = dev >= 0 && dev < PEM_MAX_DEVICES;
        if (!tracked || nonportable.granted[dev].load(std::memory_order_relaxed) == 0) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (e != cudaSuccess) return (int)e;
            if (tracked) nonportable.granted[dev].store(1, std::memory_order_relaxed);
        }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * C);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, xyz, N, M, skip, idx);
    if (e != cudaSuccess) return (int)e;
    return pem_launch_status();
}

// ------------------------------------------------------------------------------------
// gather / group
// ------------------------------------------------------------------------------------
// gather_points: the index list of a cloud is shared by its C channel rows, and every gathered float costs a
// 32-byte sector of the (b,c) row, so the kernel is bound by memory-level parallelism, not bytes.  A thread owns
// four consecutive samples (index quad read once, as int4 when aligned) and walks kGaCh channel rows with all
// 4 x kGaUnroll gathers of an unrolled step in flight before the first 16-byte store.
constexpr int kGaCh = 16;      // channel rows per CTA (grid.y = ceil(C / kGaCh))
constexpr int kGaUnroll = 4;   // rows per unrolled step: 16 independent gathers per thread
__global__ void __launch_bounds__(256)
gather_kernel(const float *__restrict__ feat, const int32_t *__restrict__ idx, int C, int N, int M,
              float *__restrict__ out, int vec) {
    const int b = blockIdx.z, c0 = blockIdx.y * kGaCh, c1 = min(C, c0 + kGaCh);
    const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (j0 >= M) return;
    const int32_t *ib = idx + (size_t)b * M + j0;
    int a[4];
    if (vec) {
        const int4 q = __ldg(reinterpret_cast<const int4 *>(ib));
        a[0] = q.x, a[1] = q.y, a[2] = q.z, a[3] = q.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = j0 + k < M ? __ldg(ib + k) : -1;
    }
    bool ok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        ok[k] = a[k] >= 0 && a[k] < N;   // out-of-range index -> 0.0f (sampling.cpp:23-44)
        if (!ok[k]) a[k] = 0;
    }
    const float *src = feat + ((size_t)b * C + c0) * N;
    float *dst = out + ((size_t)b * C + c0) * M + j0;
    int c = c0;
    for (; c + kGaUnroll <= c1; c += kGaUnroll, src += (size_t)kGaUnroll * N, dst += (size_t)kGaUnroll * M) {
        float v[kGaUnroll][4];
#pragma unroll
        for (int u = 0; u < kGaUnroll; ++u)
#pragma unroll
            for (int k = 0; k < 4; ++k) v[u][k] = __ldg(src + (size_t)u * N + a[k]);
#pragma unroll
        for (int u = 0; u < kGaUnroll; ++u) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (!ok[k]) v[u][k] = 0.f;
            float *d = dst + (size_t)u * M;
            if (vec) {
                __stcs(reinterpret_cast<float4 *>(d), make_float4(v[u][0], v[u][1], v[u][2], v[u][3]));
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (j0 + k < M) d[k] = v[u][k];
            }
        }
    }
    for (; c < c1; ++c, src += N, dst += M) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (j0 + k < M) dst[k] = ok[k] ? __ldg(src + a[k]) : 0.f;
    }
}

// Row gather on point-major data: one warp copies one C-float row.
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float *__restrict__ feat, const int32_t *__restrict__ idx, int BM, int N,
                   int C, int M, long long batch_stride, long long row_stride,
                   float *__restrict__ out, long long out_batch_stride, int vec4) {
    const int row = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (row >= BM) return;
    const int b = row / M;
    const int a = __ldg(idx + row);
    const bool ok = a >= 0 && a < N;
    const float *src = feat + (size_t)b * batch_stride + (size_t)(ok ? a : 0) * row_stride;
    float *dst = out + (size_t)b * out_batch_stride + (size_t)(row - b * M) * C;
    if (vec4) {
        const float4 *s4 = reinterpret_cast<const float4 *>(src);
        float4 *d4 = reinterpret_cast<float4 *>(dst);
        for (int c = lane_id(); c < (C >> 2); c += 32)
            d4[c] = ok ? __ldg(s4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        for (int c = lane_id(); c < C; c += 32) dst[c] = ok ? __ldg(src + c) : 0.0f;
    }
}

// out[b,c,e] = feat[b,c,idx[b,e]] over the flattened e = j*ns + k.
// Every gathered float would cost a 32-byte sector on the L2 -> SM path (8x the useful bytes; at C = 256 that
// path, not HBM, set the pace), so the R channel rows a CTA owns are first copied into shared memory (coalesced),
// the gathers become 4-byte LDS, and the index quad a thread reads (int4) serves all R rows; outputs leave as
// streaming 16-byte stores.  grid = (chunks of the E outputs, groups of R channels, B).
template <int R>
__global__ void __launch_bounds__(512)
group_smem_kernel(const float *__restrict__ feat, const int32_t *__restrict__ idx, int C, int N,
                  long long E, long long chunk, float *__restrict__ out, int vec) {
    extern __shared__ float rows[];   // R x N
    const int b = blockIdx.z, c0 = blockIdx.y * R;
    const int nr = min(R, C - c0);
    const float *src = feat + ((size_t)b * C + c0) * N;
    for (int i = threadIdx.x; i < nr * N; i += blockDim.x) rows[i] = __ldg(src + i);
    __syncthreads();
    const long long e_lo = (long long)blockIdx.x * chunk, e_hi = min(E, e_lo + chunk);
    const int32_t *ib = idx + (size_t)b * E;
    float *dst = out + ((size_t)b * C + c0) * E;
    for (long long e = e_lo + (long long)threadIdx.x * 4; e < e_hi; e += (long long)blockDim.x * 4) {
        int a[4];
        if (vec) {
            const int4 q = __ldg(reinterpret_cast<const int4 *>(ib + e));
            a[0] = q.x, a[1] = q.y, a[2] = q.z, a[3] = q.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) a[k] = e + k < e_hi ? __ldg(ib + e + k) : -1;
        }
        bool ok[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ok[k] = a[k] >= 0 && a[k] < N;   // invalid index -> 0.0f (group_points.cpp:20-45)
            if (!ok[k]) a[k] = 0;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r < nr) {
                float v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = ok[k] ? rows[r * N + a[k]] : 0.f;
                float *d = dst + (size_t)r * E + e;
                if (vec) {
                    __stcs(reinterpret_cast<float4 *>(d), make_float4(v[0], v[1], v[2], v[3]));
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (e + k < e_hi) d[k] = v[k];
                }
            }
        }
    }
}

// rows too long for shared memory: direct gathers
__global__ void __launch_bounds__(256)
group_kernel(const float *__restrict__ feat, const int32_t *__restrict__ idx, int C, int N,
             long long E, float *__restrict__ out) {
    const int b = blockIdx.z, c = blockIdx.y;
    const float *src = feat + ((size_t)b * C + c) * N;
    const int32_t *ib = idx + (size_t)b * E;
    float *dst = out + ((size_t)b * C + c) * E;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E;
         e += (long long)gridDim.x * blockDim.x) {
        const int a = __ldg(ib + e);
        dst[e] = (a >= 0 && a < N) ? __ldg(src + a) : 0.0f;
    }
}

// returns PEM_E_UNSUPPORTED when staging would not pay (few outputs per staged row): the caller gathers directly
template <int R>
static int launch_group_smem(const float *feat, const int32_t *idx, int B, int C, int N, long long E, float *out,
                             cudaStream_t st) {
    const size_t smem = (size_t)R * N * sizeof(float);
    const int groups = (C + R - 1) / R;
    // enough CTAs to fill the chip four times over, in whole 2048-output blocks
    long long gx = max(1LL, (4LL * pem_num_sms() + (long long)B * groups - 1) / ((long long)B * groups));
    long long chunk = ((E + gx - 1) / gx + 2047) / 2048 * 2048;
    chunk = min(E, chunk);
    if (chunk < 2LL * N) return PEM_E_UNSUPPORTED;  // a staged row must serve at least two outputs per element
    gx = (E + chunk - 1) / chunk;
    static PemSmemCache cached(48 * 1024);
    if (int rc = pem_ensure_dyn_smem(group_smem_kernel<R>, smem, &cached)) return rc;
    const int vec = (E % 4 == 0) && (chunk % 4 == 0) && ((uintptr_t)idx % 16 == 0) && ((uintptr_t)out % 16 == 0);
    group_smem_kernel<R><<<dim3((unsigned)gx, groups, B), 512, smem, st>>>(feat, idx, C, N, E, chunk, out, vec);
    return pem_launch_status();
}

// ------------------------------------------------------------------------------------
// ball query / fused QueryAndGroup
// ------------------------------------------------------------------------------------
constexpr int kBqCentres = 128;  // centres (= threads) per CTA
constexpr int kBqTile = 512;     // points per shared-memory tile (float4 each)

// Neighbour scan shared by ball_query / QueryAndGroup / the fused PE grouping.
// One thread = one centre.  Points stream through a shared-memory tile as float4 (one
// LDS.128 broadcast per point per warp).  For every block of 32 points the thread builds a
// 32-bit hit mask per radius (8 unfused FP ops + compare + predicated OR per point -- the
// append logic is NOT in the inner loop), then extracts the set bits in index order, which
// is exactly the reference's "first nsample hits in index order" rule.  The tile is padded
// with far-away points so the inner loop needs no bounds checks.
template <int NSCALE, typename RowT>
struct NeighbourScan {
    int cnt[NSCALE], first[NSCALE];
    RowT *row[NSCALE];

    __device__ __forceinline__ void extract(unsigned m, int s, int ns, int base) {
        while (m != 0u && cnt[s] < ns) {
            const int bit = __ffs(m) - 1;
            m &= m - 1u;
            if (cnt[s] == 0) first[s] = base + bit;
            row[s][cnt[s]++] = (RowT)(base + bit);
        }
    }

    // centres_b: (M,3); centre = centres_b[j] + eps (eps = 0 for a plain ball query)
    __device__ __forceinline__ void run(const float *__restrict__ centres_b,
                                        const float *__restrict__ xyz_b, int M, int N, float eps,
                                        const float (&r2)[NSCALE], const int (&ns)[NSCALE], int j0,
                                        float4 *s_pts) {
        const int tid = threadIdx.x, j = j0 + tid;
        const bool active = j < M;
        float qx = 0.f, qy = 0.f, qz = 0.f;
        if (active) {
            qx = __fadd_rn(centres_b[3 * j], eps);
            qy = __fadd_rn(centres_b[3 * j + 1], eps);
            qz = __fadd_rn(centres_b[3 * j + 2], eps);
        }
#pragma unroll
        for (int s = 0; s < NSCALE; ++s) { cnt[s] = active ? 0 : ns[s]; first[s] = 0; }
        for (int base = 0; base < N; base += kBqTile) {
            bool done = true;
#pragma unroll
            for (int s = 0; s < NSCALE; ++s) done = done && cnt[s] >= ns[s];
            if (__syncthreads_and(done)) break;  // also protects the tile from being overwritten
            const int nt = min(kBqTile, N - base);
            const int ntp = (nt + 31) & ~31;
            for (int k = tid; k < ntp; k += kBqCentres) {
                if (k < nt) {
                    const float *p = xyz_b + 3 * (size_t)(base + k);
                    s_pts[k] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
                } else {
                    s_pts[k] = make_float4(3e18f, 3e18f, 3e18f, 0.f);  // d2 = +inf: never a hit
                }
            }
            __syncthreads();
            if (__all_sync(kFull, done)) continue;
            for (int kb = 0; kb < ntp; kb += 32) {
                unsigned m[NSCALE];
#pragma unroll
                for (int s = 0; s < NSCALE; ++s) m[s] = 0u;
#pragma unroll
                for (int u = 0; u < 32; ++u) {
                    const float4 p = s_pts[kb + u];
                    const float d2 = sqdist_unfused(qx, qy, qz, p.x, p.y, p.z);
#pragma unroll
                    for (int s = 0; s < NSCALE; ++s)
                        if (d2 < r2[s]) m[s] |= 1u << u;
                }
#pragma unroll
                for (int s = 0; s < NSCALE; ++s) extract(m[s], s, ns[s], base + kb);
            }
        }
        if (active) {
#pragma unroll
            for (int s = 0; s < NSCALE; ++s)
                for (int l = cnt[s]; l < ns[s]; ++l) row[s][l] = (RowT)first[s];  // back-fill; no hit -> 0
        }
        __syncthreads();
    }
};

// Fills s_rows[tid*(ns+1) + l], l in [0,ns), with the neighbour list of centre j0+tid.
__device__ __forceinline__ void ball_query_rows(const float *__restrict__ centres_b,
                                                const float *__restrict__ xyz_b, int M, int N,
                                                float r2, int ns, int j0, float4 *s_pts,
                                                int *s_rows) {
    NeighbourScan<1, int> scan;
    scan.row[0] = s_rows + threadIdx.x * (ns + 1);
    const float r2a[1] = {r2};
    const int nsa[1] = {ns};
    scan.run(centres_b, xyz_b, M, N, 0.f, r2a, nsa, j0, s_pts);
}

__global__ void __launch_bounds__(kBqCentres)
ball_query_kernel(const float *__restrict__ centres, const float *__restrict__ xyz, int M, int N,
                  float r2, int ns, int32_t *__restrict__ idx) {
    extern __shared__ float4 s_dyn[];
    float4 *s_pts = s_dyn;
    int *s_rows = reinterpret_cast<int *>(s_dyn + kBqTile);
    const int b = blockIdx.y, j0 = blockIdx.x * kBqCentres;
    ball_query_rows(centres + (size_t)b * M * 3, xyz + (size_t)b * N * 3, M, N, r2, ns, j0, s_pts,
                    s_rows);
    const int rows = min(kBqCentres, M - j0);
    int32_t *dst = idx + ((size_t)b * M + j0) * ns;
    for (int e = threadIdx.x; e < rows * ns; e += kBqCentres) {
        const int jj = e / ns, l = e - jj * ns;
        dst[e] = s_rows[jj * (ns + 1) + l];
    }
}

__global__ void __launch_bounds__(kBqCentres)
query_and_group_kernel(const float *__restrict__ xyz, const float *__restrict__ centres,
                       const float *__restrict__ feat, int N, int M, int C, float r2, int ns,
                       float *__restrict__ out, int32_t *__restrict__ idx_out) {
    extern __shared__ float4 s_dyn[];
    float4 *s_pts = s_dyn;
    int *s_rows = reinterpret_cast<int *>(s_dyn + kBqTile);
    const int b = blockIdx.y, j0 = blockIdx.x * kBqCentres;
    const float *xyz_b = xyz + (size_t)b * N * 3;
    const float *cen_b = centres + (size_t)b * M * 3;
    ball_query_rows(cen_b, xyz_b, M, N, r2, ns, j0, s_pts, s_rows);
    const int rows = min(kBqCentres, M - j0);
    const int CT = 3 + C;
    const size_t plane = (size_t)M * ns;
    float *out_b = out + (size_t)b * CT * plane + (size_t)j0 * ns;
    for (int e = threadIdx.x; e < rows * ns; e += kBqCentres) {
        const int jj = e / ns, l = e - jj * ns;
        const int a = s_rows[jj * (ns + 1) + l];
        if (idx_out) idx_out[((size_t)b * M + j0) * ns + e] = a;
        const float *p = xyz_b + 3 * (size_t)a;
        const float *q = cen_b + 3 * (size_t)(j0 + jj);
        // grouped_xyz -= new_xyz (pointnet2_utils.py:376): plain fp32 subtraction
        out_b[e] = __fsub_rn(__ldg(p), __ldg(q));
        out_b[plane + e] = __fsub_rn(__ldg(p + 1), __ldg(q + 1));
        out_b[2 * plane + e] = __fsub_rn(__ldg(p + 2), __ldg(q + 2));
        for (int c = 0; c < C; ++c)
            out_b[(size_t)(3 + c) * plane + e] = __ldg(feat + ((size_t)b * C + c) * N + a);
    }
}

// out[b,i,:] = (p[b,i,:] - t[b]) @ R[b]  (+ eps on every coordinate): the point transform of
// fine_point_matching.py:45 and the `pts + 1e-8` query centres of fine_point_matching.py:117.
__global__ void __launch_bounds__(256)
transform_points_kernel(const float *__restrict__ p, const float *__restrict__ R,
                        const float *__restrict__ t, int N, float eps, float *__restrict__ out) {
    const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float *q = p + ((size_t)b * N + i) * 3;
    float v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2);
    float o0 = v0, o1 = v1, o2 = v2;
    if (R != nullptr) {
        const float *Rb = R + (size_t)b * 9, *tb = t + (size_t)b * 3;
        v0 -= tb[0]; v1 -= tb[1]; v2 -= tb[2];
        o0 = v0 * Rb[0] + v1 * Rb[3] + v2 * Rb[6];
        o1 = v0 * Rb[1] + v1 * Rb[4] + v2 * Rb[7];
        o2 = v0 * Rb[2] + v1 * Rb[5] + v2 * Rb[8];
    }
    float *o = out + ((size_t)b * N + i) * 3;
    o[0] = __fadd_rn(o0, eps); o[1] = __fadd_rn(o1, eps); o[2] = __fadd_rn(o2, eps);
}

// Fused PositionalEncoding front end (fine_point_matching.py:113-137): both QueryAndGroup
// scales of one cloud in a single pass over the points (the distance is computed once and
// compared against both radii).  centre = xyz + eps; channels 0-2 = neighbour - centre,
// channels 3-5 = neighbour (the "features" of PE are the coordinates themselves).
//
// Parallel decomposition: with 16 clouds x 2048 centres a thread-per-centre scan exposes only
// 32k threads, far too few for 148 SMs.  So each centre is scanned by kPeSeg threads; thread
// (centre, seg) owns a contiguous quarter of the current 1024-point tile and keeps its hit
// masks (one 32-bit word per 32 points per radius) in registers.  Segment hit counts go through
// shared memory; each thread then knows where its hits start in the centre's neighbour row
// (hits of earlier tiles + hits of lower segments) and appends them in index order -- exactly
// the reference's "first nsample hits in index order".  A warp holds 32 consecutive centres
// of ONE segment, so every point is a single LDS.128 broadcast.
constexpr int kPeCentres = 64;   // centres per CTA
constexpr int kPeSeg = 4;        // threads per centre
constexpr int kPeThreads = kPeCentres * kPeSeg;
constexpr int kPeTile = 1024;    // points per tile
constexpr int kPeSegWords = kPeTile / 32 / kPeSeg;  // mask words per thread per radius

// CLOUD_IN_SMEM (N <= 4096): the whole cloud, padded to a multiple of kPeTile, is staged once as
// float4 and serves both the scan and the final gather: a random 16-byte shared-memory read per
// neighbour instead of three fully divergent global loads (32 L1 wavefronts each).
//
// MASKED: instead of testing every point, the scan visits only the set bits of a per-centre
// candidate mask (B,N,ceil(N/32)) produced by the tensor-core filter (pe_mask_tc.cu) for the LARGER
// radius -- possibly on the cloud as it was BEFORE a rigid motion, since ball membership survives
// one up to rounding and the filter is generous by more than that.  Every candidate still goes
// through the exact test below on THIS cloud, so the outputs do not change.
template <bool CLOUD_IN_SMEM, bool MASKED>
__global__ void __launch_bounds__(kPeThreads)
pe_group2_kernel(const float *__restrict__ xyz, int N, float eps, float r2a, int ns1,
                 float *__restrict__ out1, float r2b, int ns2, float *__restrict__ out2,
                 const uint32_t *__restrict__ cmask, int W, int32_t *__restrict__ idx1 /* (B,N,ns1) or null */,
                 int32_t *__restrict__ idx2 /* (B,N,ns2) or null */) {
    extern __shared__ float4 s_dyn[];
    const int n_pts_smem = CLOUD_IN_SMEM ? (N + kPeTile - 1) / kPeTile * kPeTile : kPeTile;
    float4 *s_pts = s_dyn;                                                   // n_pts_smem
    unsigned short *rows1 = reinterpret_cast<unsigned short *>(s_dyn + n_pts_smem);
    const int st1 = ns1 + 2, st2 = ns2 + 2;                                  // odd strides in 32-bit words
    unsigned short *rows2 = rows1 + kPeCentres * st1;
    __shared__ int s_seg[2][kPeSeg][kPeCentres];  // hits of (radius, segment, centre) in this tile
    const int tid = threadIdx.x, b = blockIdx.y, j0 = blockIdx.x * kPeCentres;
    const int c = tid & (kPeCentres - 1), seg = tid / kPeCentres;  // seg is warp-uniform
    const float *xyz_b = xyz + (size_t)b * N * 3;
    const int j = j0 + c;
    const bool active = j < N;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        qx = __fadd_rn(xyz_b[3 * j], eps);
        qy = __fadd_rn(xyz_b[3 * j + 1], eps);
        qz = __fadd_rn(xyz_b[3 * j + 2], eps);
    }
    unsigned short *row1 = rows1 + c * st1, *row2 = rows2 + c * st2;
    int cnt1 = active ? 0 : ns1, cnt2 = active ? 0 : ns2;  // identical in the 4 threads of a centre
    if (CLOUD_IN_SMEM) {
        for (int k = tid; k < n_pts_smem; k += kPeThreads) {
            if (k < N) {
                const float *p = xyz_b + 3 * (size_t)k;
                s_pts[k] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
            } else {
                s_pts[k] = make_float4(3e18f, 3e18f, 3e18f, 0.f);  // d2 = +inf: never a hit
            }
        }
    }
    for (int base = 0; base < N; base += kPeTile) {
        if (__syncthreads_and(cnt1 >= ns1 && cnt2 >= ns2)) break;  // every list of this CTA is full
        if (!CLOUD_IN_SMEM) {
            const int nt = min(kPeTile, N - base);
            for (int k = tid; k < kPeTile; k += kPeThreads) {
                if (k < nt) {
                    const float *p = xyz_b + 3 * (size_t)(base + k);
                    s_pts[k] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
                } else {
                    s_pts[k] = make_float4(3e18f, 3e18f, 3e18f, 0.f);
                }
            }
            __syncthreads();
        }
        const float4 *tile = CLOUD_IN_SMEM ? s_pts + base : s_pts;
        unsigned m1[kPeSegWords], m2[kPeSegWords];
        int h1 = 0, h2 = 0;
        unsigned cw[kPeSegWords];  // candidate words of this thread's quarter of the tile
        if (MASKED) {
            const int wi0 = (base >> 5) + seg * kPeSegWords;
            const uint32_t *mrow = cmask + ((size_t)b * N + (active ? j : 0)) * W;
#pragma unroll
            for (int w = 0; w < kPeSegWords; ++w) cw[w] = 0u;
            if (active) {
                if ((W & 3) == 0 && wi0 + kPeSegWords <= W) {
#pragma unroll
                    for (int w = 0; w < kPeSegWords; w += 4) {
                        const uint4 q4 = __ldg(reinterpret_cast<const uint4 *>(mrow + wi0 + w));
                        cw[w] = q4.x; cw[w + 1] = q4.y; cw[w + 2] = q4.z; cw[w + 3] = q4.w;
                    }
                } else {
#pragma unroll
                    for (int w = 0; w < kPeSegWords; ++w)
                        if (wi0 + w < W) cw[w] = __ldg(mrow + wi0 + w);
                }
            }
        }
#pragma unroll
        for (int w = 0; w < kPeSegWords; ++w) {
            const float4 *pp = tile + (seg * kPeSegWords + w) * 32;
            unsigned a1 = 0u, a2 = 0u;
            if (MASKED) {
                unsigned cm = cw[w];
                while (cm != 0u) {
                    const int u = __ffs(cm) - 1;
                    cm &= cm - 1u;
                    const float4 p = pp[u];
                    const float d2 = sqdist_unfused(qx, qy, qz, p.x, p.y, p.z);
                    if (d2 < r2a) a1 |= 1u << u;
                    if (d2 < r2b) a2 |= 1u << u;
                }
            } else {
#pragma unroll
                for (int u = 0; u < 32; ++u) {
                    const float4 p = pp[u];
                    const float d2 = sqdist_unfused(qx, qy, qz, p.x, p.y, p.z);
                    if (d2 < r2a) a1 |= 1u << u;
                    if (d2 < r2b) a2 |= 1u << u;
                }
            }
            m1[w] = a1; m2[w] = a2;
            h1 += __popc(a1); h2 += __popc(a2);
        }
        s_seg[0][seg][c] = h1;
        s_seg[1][seg][c] = h2;
        __syncthreads();
        // where this thread's hits start, and the centre's new totals
        int off1 = cnt1, off2 = cnt2, tot1 = cnt1, tot2 = cnt2;
#pragma unroll
        for (int sg = 0; sg < kPeSeg; ++sg) {
            const int a1 = s_seg[0][sg][c], a2 = s_seg[1][sg][c];
            if (sg < seg) { off1 += a1; off2 += a2; }
            tot1 += a1; tot2 += a2;
        }
        const int id0 = base + seg * kPeSegWords * 32;
#pragma unroll
        for (int w = 0; w < kPeSegWords; ++w) {
            unsigned m = m1[w];
            while (m != 0u && off1 < ns1) {
                const int bit = __ffs(m) - 1;
                m &= m - 1u;
                row1[off1++] = (unsigned short)(id0 + w * 32 + bit);
            }
            m = m2[w];
            while (m != 0u && off2 < ns2) {
                const int bit = __ffs(m) - 1;
                m &= m - 1u;
                row2[off2++] = (unsigned short)(id0 + w * 32 + bit);
            }
        }
        cnt1 = min(tot1, ns1);
        cnt2 = min(tot2, ns2);
    }
    __syncthreads();
    if (active && seg == 0) {  // back-fill with the first hit; no hit -> zeros
        const unsigned short f1 = cnt1 ? row1[0] : (unsigned short)0, f2 = cnt2 ? row2[0] : (unsigned short)0;
        for (int l = cnt1; l < ns1; ++l) row1[l] = f1;
        for (int l = cnt2; l < ns2; ++l) row2[l] = f2;
    }
    __syncthreads();
    const int rows = min(kPeCentres, N - j0);
#pragma unroll 1
    for (int s = 0; s < 2; ++s) {
        const int nss = s ? ns2 : ns1, stride = s ? st2 : st1;
        const unsigned short *rw = s ? rows2 : rows1;
        if (int32_t *io = s ? idx2 : idx1) {  // the neighbour lists themselves (ball_query's output)
            int32_t *ib = io + ((size_t)b * N + j0) * nss;
            for (int e = tid; e < rows * nss; e += kPeThreads) ib[e] = rw[(e / nss) * stride + e % nss];
        }
        if ((s ? out2 : out1) == nullptr) continue;
        const size_t plane = (size_t)N * nss;
        float *out_b = (s ? out2 : out1) + (size_t)b * 6 * plane + (size_t)j0 * nss;
        const bool pow2 = (nss & (nss - 1)) == 0;
        const int sh = 31 - __clz(nss);
        for (int e = tid; e < rows * nss; e += kPeThreads) {
            const int jj = pow2 ? (e >> sh) : (e / nss), l = e - jj * nss;
            const int a = rw[jj * stride + l];
            float px, py, pz, cx, cy, cz;
            if (CLOUD_IN_SMEM) {
                const float4 p = s_pts[a], q = s_pts[j0 + jj];
                px = p.x; py = p.y; pz = p.z; cx = q.x; cy = q.y; cz = q.z;
            } else {
                const float *p = xyz_b + 3 * (size_t)a;
                const float *q = xyz_b + 3 * (size_t)(j0 + jj);
                px = __ldg(p); py = __ldg(p + 1); pz = __ldg(p + 2);
                cx = __ldg(q); cy = __ldg(q + 1); cz = __ldg(q + 2);
            }
            out_b[e] = __fsub_rn(px, __fadd_rn(cx, eps));
            out_b[plane + e] = __fsub_rn(py, __fadd_rn(cy, eps));
            out_b[2 * plane + e] = __fsub_rn(pz, __fadd_rn(cz, eps));
            out_b[3 * plane + e] = px;
            out_b[4 * plane + e] = py;
            out_b[5 * plane + e] = pz;
        }
    }
}

// ------------------------------------------------------------------------------------
// Masked PE grouping, one WARP per centre (round 2).  pe_group2_kernel<.., MASKED> walks the candidate
// masks with one thread per (centre, quarter tile): its bit loops diverge over 32 different centres and
// every tile costs two CTA barriers, ~500 warp instructions per centre.  Here a warp owns a centre: lane l
// takes the mask words of columns [l * wpl * 32, (l + 1) * wpl * 32) (one coalesced 256-byte read per centre
// at N = 2048), re-tests its set bits exactly, and the hits are compacted in index order by one packed
// warp scan; the neighbour rows then leave with the lanes along the neighbour axis (128-byte stores).
// Same arithmetic, same "first nsample hits in index order, back-fill with the first, none -> 0" rule.
// ------------------------------------------------------------------------------------
constexpr int kPwWarps = 8;       // warps per CTA
constexpr int kPwPerWarp = 8;     // centres per warp
constexpr int kPwCentres = kPwWarps * kPwPerWarp;
constexpr int kPwCand = 256;      // candidates per round of a centre (index-ordered list in shared memory)
// kPwMaxWpl = mask words per lane (1, 2 or 4: N <= 1024, 2048, 4096), a template argument so that the
// per-lane word arrays stay as small as the cloud needs: <= 51 registers keep five CTAs on an SM and the
// 512 CTAs of 16 clouds x 2048 centres in ONE wave (at 64 registers they took two)
// NS1 / NS2: the neighbour counts as compile-time constants (0 = take the run-time arguments): the PEM pair
// (32, 64) unrolls the output loops and keeps the plane offsets in immediates.
template <int kPwMaxWpl, int NS1, int NS2>
__global__ void __launch_bounds__(32 * kPwWarps, kPwMaxWpl <= 2 ? 5 : 4)
pe_group2_warp_kernel(const float *__restrict__ xyz, int N, float eps, float r2a, int ns1_rt,
                      float *__restrict__ out1, float r2b, int ns2_rt, float *__restrict__ out2,
                      const uint32_t *__restrict__ cmask, int W, int32_t *__restrict__ idx1,
                      int32_t *__restrict__ idx2) {
    extern __shared__ float4 s_dyn[];
    const int ns1 = NS1 ? NS1 : ns1_rt, ns2 = NS2 ? NS2 : ns2_rt;
    // the cloud, one spare entry after every lane's share of 32 * wpl points: lane l's candidates then start at
    // bank group (l + bit) mod 8 instead of all lanes in one (their shares are a multiple of 1 KB apart).
    // Candidate list and neighbour rows hold these PADDED positions (column + owning lane).
    constexpr int lg = kPwMaxWpl == 1 ? 0 : kPwMaxWpl == 2 ? 1 : 2, wpl = kPwMaxWpl;  // mask words per lane
    constexpr int share = 32 * wpl;
    auto P = [&](int k_) { return k_ + (k_ >> (5 + lg)); };
    auto unP = [&](int pk) { return pk - pk / (share + 1); };   // (only where indices leave the kernel)
    float4 *s_pts = s_dyn;
    unsigned short *rows = reinterpret_cast<unsigned short *>(s_dyn + P(N) + 1);
    const int warp = warp_id(), lane = lane_id(), b = blockIdx.y;
    const int row_len = ns1 + ns2;
    unsigned short *row1 = rows + warp * row_len, *row2 = row1 + ns1;
    unsigned short *cand = rows + kPwWarps * row_len + warp * kPwCand;
    const float *xyz_b = xyz + (size_t)b * N * 3;
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        const float *p = xyz_b + 3 * (size_t)k;
        s_pts[P(k)] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
    }
    __syncthreads();
    const int j_begin = blockIdx.x * kPwCentres + warp * kPwPerWarp;
    const int j_end = min(N, j_begin + kPwPerWarp);
    const unsigned below = (1u << lane) - 1u;
    const int pid0 = lane * share + lane;   // padded position of this lane's first column
    // the candidate words of the warp's first centre are in flight before the loop; inside it the next
    // centre's words are fetched before the current one is processed
    uint32_t cw_next[kPwMaxWpl];
    auto fetch = [&](int j) {
        const uint32_t *mrow = cmask + ((size_t)b * N + j) * W + lane * wpl;
#pragma unroll
        for (int w = 0; w < kPwMaxWpl; ++w) cw_next[w] = (lane * wpl + w < W) ? __ldg(mrow + w) : 0u;
    };
    if (j_begin < j_end) fetch(j_begin);
    for (int j = j_begin; j < j_end; ++j) {
        uint32_t cw[kPwMaxWpl];
#pragma unroll
        for (int w = 0; w < kPwMaxWpl; ++w) cw[w] = cw_next[w];
        if (j + 1 < j_end) fetch(j + 1);
        const float4 c = s_pts[P(j)];
        const float qx = __fadd_rn(c.x, eps), qy = __fadd_rn(c.y, eps), qz = __fadd_rn(c.z, eps);
        // Lanes = candidates.  The set bits are first laid out as an index-ordered list (a lane's share of the
        // columns is contiguous, so one scan over the lanes' counts orders them), then 32 candidates at a time
        // go through the exact test; the hit ballots are already in index order, so a hit's slot is the running
        // total plus the hits below its lane.  (The per-lane bit loops that remain only COPY positions: ~2 per
        // lane; walking the masks with one lane per centre-share cost 760 warp instructions per centre.)
        int n = 0;
#pragma unroll
        for (int w = 0; w < kPwMaxWpl; ++w) n += __popc(cw[w]);
        int incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += o;
        }
        const int ncand = __shfl_sync(kFull, incl, 31), first_rank = incl - n;
        int base1 = 0, base2 = 0;
        for (int lo = 0; lo < ncand && (base1 < ns1 || base2 < ns2); lo += kPwCand) {
            if (ncand <= kPwCand) {  // the whole list at once (the rule)
                unsigned short *dst = cand + first_rank;
#pragma unroll
                for (int w = 0; w < kPwMaxWpl; ++w) {
                    uint32_t cm = cw[w];
                    while (cm != 0u) {
                        *dst++ = (unsigned short)(pid0 + w * 32 + __ffs(cm) - 1);
                        cm &= cm - 1u;
                    }
                }
            } else if (first_rank < lo + kPwCand && first_rank + n > lo) {  // ranks [lo, lo + kPwCand)
                int rank = first_rank;
#pragma unroll
                for (int w = 0; w < kPwMaxWpl; ++w) {
                    uint32_t cm = cw[w];
                    while (cm != 0u) {
                        const int u = __ffs(cm) - 1;
                        cm &= cm - 1u;
                        if (rank >= lo && rank < lo + kPwCand) cand[rank - lo] = (unsigned short)(pid0 + w * 32 + u);
                        ++rank;
                    }
                }
            }
            __syncwarp();
            const int cnt = min(kPwCand, ncand - lo);
            for (int c0 = 0; c0 < cnt && (base1 < ns1 || base2 < ns2); c0 += 32) {
                const int i = c0 + lane;
                const bool valid = i < cnt;
                const int pc = valid ? cand[i] : 0;
                const float4 p = s_pts[pc];
                const float d2 = sqdist_unfused(qx, qy, qz, p.x, p.y, p.z);
                const bool hit1 = valid && d2 < r2a, hit2 = valid && d2 < r2b;
                const unsigned m1 = __ballot_sync(kFull, hit1), m2 = __ballot_sync(kFull, hit2);
                const int r1 = base1 + __popc(m1 & below), r2 = base2 + __popc(m2 & below);
                if (hit1 && r1 < ns1) row1[r1] = (unsigned short)pc;
                if (hit2 && r2 < ns2) row2[r2] = (unsigned short)pc;
                base1 += __popc(m1);
                base2 += __popc(m2);
            }
            __syncwarp();  // the list is rewritten by the next round
        }
        const int tot1 = min(base1, ns1), tot2 = min(base2, ns2);
        // back-fill with the first hit; no hit -> point 0 (padded position 0)
        {
            const unsigned short f1 = tot1 ? row1[0] : (unsigned short)0, f2 = tot2 ? row2[0] : (unsigned short)0;
            for (int l = tot1 + lane; l < ns1; l += 32) row1[l] = f1;
            for (int l = tot2 + lane; l < ns2; l += 32) row2[l] = f2;
        }
        __syncwarp();
        // outputs: lanes along the neighbour axis
        auto emit = [&](const unsigned short *rw, int nss, float *ob, int32_t *io) {
            if (io) {
                int32_t *ib = io + ((size_t)b * N + j) * nss;
                for (int l = lane; l < nss; l += 32) ib[l] = unP(rw[l]);
            }
            if (ob == nullptr) return;
            const int plane = N * nss;   // (N <= 4096, nss <= 128: 6 planes fit 32-bit offsets)
            float *o = ob + (size_t)b * 6 * plane + (size_t)j * nss + lane;
#pragma unroll
            for (int l = 0; l < 128; l += 32) {   // nsample <= 128
                if (l + lane >= nss) break;
                const float4 p = s_pts[rw[l + lane]];
                o[l] = __fsub_rn(p.x, qx);
                o[plane + l] = __fsub_rn(p.y, qy);
                o[2 * plane + l] = __fsub_rn(p.z, qz);
                o[3 * plane + l] = p.x;
                o[4 * plane + l] = p.y;
                o[5 * plane + l] = p.z;
            }
        };
        emit(row1, ns1, out1, idx1);
        emit(row2, ns2, out2, idx2);
        __syncwarp();  // the rows are rewritten by the next centre
    }
}

static size_t bq_smem_bytes(int ns) {
    return (size_t)kBqTile * sizeof(float4) + (size_t)kBqCentres * (ns + 1) * sizeof(int);
}

}  // namespace pem

// ------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------
using namespace pem;

#ifdef PEM_DEV_KNOBS   // make DEV=1 (tools/fps_variants.py): alternative launch shapes of the sampling kernel
static int g_fps_variant = 0;
extern "C" void pem_dev_set_fps_variant(int v) { g_fps_variant = v; }
#else
static constexpr int g_fps_variant = 0;
#endif

extern "C" size_t pem_fps_workspace_bytes(int B, int N) {
    if (B <= 0 || N <= 16 * 16384) return 0;
    return (size_t)B * N * sizeof(float);
}

extern "C" int pem_fps(const float *xyz, int B, int N, int M, int flags, int32_t *idx,
                       void *workspace, size_t workspace_bytes, pem_stream_t stream) {
    PEM_REQUIRE_PTR(xyz);
    PEM_REQUIRE_PTR(idx);
    PEM_REQUIRE(B > 0 && N > 0 && M > 0, PEM_E_SHAPE);
    cudaStream_t st = pem_cu_stream(stream);
    const int skip = (flags & PEM_FPS_SKIP_NEAR_ORIGIN) ? 1 : 0;
    if (N <= 1024) return launch_fps_cta<128, 8, true>(xyz, B, N, M, skip, idx, st);
    if (N <= 2048) {
        switch (g_fps_variant) {  // development knob (pem_dev_set_fps_variant); 0 = shipped choice
            case 2: return launch_fps_cta<512, 4, true>(xyz, B, N, M, skip, idx, st);
            case 3: return launch_fps_cta<1024, 2, true>(xyz, B, N, M, skip, idx, st);
            case 4: return launch_fps_cta<256, 8, true>(xyz, B, N, M, skip, idx, st);
            default: return launch_fps_cta<128, 16, true>(xyz, B, N, M, skip, idx, st);  // fewest warps to synchronise
        }
    }
    if (N <= 4096) return launch_fps_cta<512, 8, true>(xyz, B, N, M, skip, idx, st);
    if (N <= 8192) return launch_fps_cta<1024, 8, true>(xyz, B, N, M, skip, idx, st);
    // larger clouds: thread-block clusters (slabs of <= 8192 points in registers), as long as the
    // clusters of all clouds fit on the chip at once; many medium clouds stay one CTA per cloud
    if (N <= 16384) {
        if (2 * B <= PEM_NUM_SMS) return launch_fps_cluster<1024, 8, true>(xyz, B, N, M, skip, idx, 2, st);
        return launch_fps_cta<1024, 16, false>(xyz, B, N, M, skip, idx, st);
    }
    if (N <= 65536) {
        const int C = N <= 32768 ? 4 : 8;
        return launch_fps_cluster<1024, 8, true>(xyz, B, N, M, skip, idx, C, st);
    }
    if (N <= 16 * 16384)  // e.g. the 210 000-point template cloud: 16 CTAs, coordinates in shared memory
        return launch_fps_cluster<1024, 16, false>(xyz, B, N, M, skip, idx, 16, st);
    PEM_REQUIRE(workspace != nullptr && workspace_bytes >= pem_fps_workspace_bytes(B, N),
                PEM_E_WORKSPACE);
    fps_stream_kernel<1024><<<B, 1024, 0, st>>>(xyz, N, M, skip, (float *)workspace, idx);
    return pem_launch_status();
}

// FPS + the gather of the sampled points (sample_pts_feats, MU:70-84, point half) in one launch where
// the cloud lives in one CTA (N <= 2048: the PEM shapes); otherwise pem_fps followed by the row gather.
extern "C" int pem_fps_gather(const float *xyz, int B, int N, int M, int flags, int32_t *idx,
                              float *new_xyz, void *workspace, size_t workspace_bytes,
                              pem_stream_t stream) {
    PEM_REQUIRE_PTR(xyz);
    PEM_REQUIRE_PTR(idx);
    PEM_REQUIRE_PTR(new_xyz);
    PEM_REQUIRE(B > 0 && N > 0 && M > 0, PEM_E_SHAPE);
    cudaStream_t st = pem_cu_stream(stream);
    const int skip = (flags & PEM_FPS_SKIP_NEAR_ORIGIN) ? 1 : 0;
    if (N <= 1024) return launch_fps_cta<128, 8, true>(xyz, B, N, M, skip, idx, st, new_xyz);
    if (N <= 2048 && g_fps_variant == 0) return launch_fps_cta<128, 16, true>(xyz, B, N, M, skip, idx, st, new_xyz);
    if (int rc = pem_fps(xyz, B, N, M, flags, idx, workspace, workspace_bytes, stream)) return rc;
    return pem_gather_rows(xyz, idx, B, N, 3, M, (long long)N * 3, 3, new_xyz, 0, stream);
}

extern "C" int pem_gather(const float *feat, const int32_t *idx, int B, int C, int N, int M,
                          float *out, pem_stream_t stream) {
    PEM_REQUIRE_PTR(feat);
    PEM_REQUIRE_PTR(idx);
    PEM_REQUIRE_PTR(out);
    PEM_REQUIRE(B > 0 && C > 0 && N > 0 && M > 0, PEM_E_SHAPE);
    PEM_REQUIRE(B <= 65535 && (C + kGaCh - 1) / kGaCh <= 65535, PEM_E_SHAPE);
    const int quads = (M + 3) / 4, threads = quads >= 256 ? 256 : (quads + 31) / 32 * 32;
    const int vec = (M % 4 == 0) && ((uintptr_t)idx % 16 == 0) && ((uintptr_t)out % 16 == 0);
    gather_kernel<<<dim3((quads + threads - 1) / threads, (C + kGaCh - 1) / kGaCh, B), threads, 0,
                    pem_cu_stream(stream)>>>(feat, idx, C, N, M, out, vec);
    return pem_launch_status();
}

extern "C" int pem_gather_rows(const float *feat, const int32_t *idx, int B, int N, int C, int M,
                               long long batch_stride, long long row_stride, float *out,
                               long long out_batch_stride, pem_stream_t stream) {
    PEM_REQUIRE_PTR(feat);
    PEM_REQUIRE_PTR(idx);
    PEM_REQUIRE_PTR(out);
    PEM_REQUIRE(B > 0 && C > 0 && N > 0 && M > 0 && row_stride >= C, PEM_E_SHAPE);
    if (out_batch_stride <= 0) out_batch_stride = (long long)M * C;
    PEM_REQUIRE(out_batch_stride >= (long long)M * C, PEM_E_SHAPE);
    const int rows = B * M;
    const int vec4 = (C % 4 == 0) && (row_stride % 4 == 0) && (batch_stride % 4 == 0) &&
                     (out_batch_stride % 4 == 0) &&
                     ((uintptr_t)feat % 16 == 0) && ((uintptr_t)out % 16 == 0);
    gather_rows_kernel<<<(rows + 7) / 8, 256, 0, pem_cu_stream(stream)>>>(
        feat, idx, rows, N, C, M, batch_stride, row_stride, out, out_batch_stride, vec4);
    return pem_launch_status();
}

extern "C" int pem_group(const float *feat, const int32_t *idx, int B, int C, int N, int M, int ns,
                         float *out, pem_stream_t stream) {
    PEM_REQUIRE_PTR(feat);
    PEM_REQUIRE_PTR(idx);
    PEM_REQUIRE_PTR(out);
    PEM_REQUIRE(B > 0 && C > 0 && N > 0 && M > 0 && ns > 0 && C <= 65535 && B <= 65535,
                PEM_E_SHAPE);
    const long long E = (long long)M * ns;
    cudaStream_t st = pem_cu_stream(stream);
    // rows staged in shared memory, three CTAs per SM (<= 72 KB each) so that one CTA's staging hides under the
    // others' gathers; as many rows per CTA as fit (the index quads are read once per CTA)
    const size_t row = (size_t)N * sizeof(float);
    int rc = PEM_E_UNSUPPORTED;
    if (row * 4 <= 72 * 1024 && C >= 4) rc = launch_group_smem<4>(feat, idx, B, C, N, E, out, st);
    else if (row * 3 <= 72 * 1024 && C >= 3) rc = launch_group_smem<3>(feat, idx, B, C, N, E, out, st);
    else if (row <= 200 * 1024) rc = launch_group_smem<1>(feat, idx, B, C, N, E, out, st);
    if (rc != PEM_E_UNSUPPORTED) return rc;
    const int gx = (int)((E + 1023) / 1024);  // 4 elements per thread
    group_kernel<<<dim3(gx, C, B), 256, 0, st>>>(feat, idx, C, N, E, out);
    return pem_launch_status();
}

extern "C" int pem_ball_query(const float *new_xyz, const float *xyz, int B, int M, int N,
                              float radius, int nsample, int32_t *idx, pem_stream_t stream) {
    PEM_REQUIRE_PTR(new_xyz);
    PEM_REQUIRE_PTR(xyz);
    PEM_REQUIRE_PTR(idx);
    PEM_REQUIRE(B > 0 && M > 0 && N > 0 && nsample > 0 && B <= 65535, PEM_E_SHAPE);
    PEM_REQUIRE(nsample <= 128, PEM_E_UNSUPPORTED);
    const size_t smem = bq_smem_bytes(nsample);
    static PemSmemCache cached(0);
    if (int rc = pem_ensure_dyn_smem(ball_query_kernel, smem, &cached)) return rc;
    const float r2 = radius * radius;  // fp32 product, as ball_query.cpp:20
    ball_query_kernel<<<dim3((M + kBqCentres - 1) / kBqCentres, B), kBqCentres, smem,
                        pem_cu_stream(stream)>>>(new_xyz, xyz, M, N, r2, nsample, idx);
    return pem_launch_status();
}

extern "C" int pem_query_and_group(const float *xyz, const float *new_xyz, const float *feat, int B,
                                   int N, int M, int C, float radius, int nsample, float *out,
                                   int32_t *idx_out, pem_stream_t stream) {
    PEM_REQUIRE_PTR(xyz);
    PEM_REQUIRE_PTR(new_xyz);
    PEM_REQUIRE_PTR(out);
    PEM_REQUIRE(B > 0 && M > 0 && N > 0 && nsample > 0 && C >= 0 && B <= 65535, PEM_E_SHAPE);
    PEM_REQUIRE(C == 0 || feat != nullptr, PEM_E_NULL);
    PEM_REQUIRE(nsample <= 128, PEM_E_UNSUPPORTED);
    const size_t smem = bq_smem_bytes(nsample);
    static PemSmemCache cached(0);
    if (int rc = pem_ensure_dyn_smem(query_and_group_kernel, smem, &cached)) return rc;
    const float r2 = radius * radius;
    query_and_group_kernel<<<dim3((M + kBqCentres - 1) / kBqCentres, B), kBqCentres, smem,
                             pem_cu_stream(stream)>>>(xyz, new_xyz, feat, N, M, C, r2, nsample,
                                                      out, idx_out);
    return pem_launch_status();
}

extern "C" int pem_transform_points(const float *pts, const float *R, const float *t, int B, int N,
                                    float eps, float *out, pem_stream_t stream) {
    PEM_REQUIRE_PTR(pts);
    PEM_REQUIRE_PTR(out);
    PEM_REQUIRE(B > 0 && N > 0 && B <= 65535, PEM_E_SHAPE);
    PEM_REQUIRE((R == nullptr) == (t == nullptr), PEM_E_NULL);
    transform_points_kernel<<<dim3((N + 255) / 256, B), 256, 0, pem_cu_stream(stream)>>>(pts, R, t, N,
                                                                                       eps, out);
    return pem_launch_status();
}

extern "C" int pem_pe_group2(const float *xyz, int B, int N, float centre_eps, float radius1,
                             int nsample1, float *out1, float radius2, int nsample2, float *out2,
                             pem_stream_t stream) {
    PEM_REQUIRE_PTR(xyz);
    PEM_REQUIRE_PTR(out1);
    PEM_REQUIRE_PTR(out2);
    PEM_REQUIRE(B > 0 && N > 0 && nsample1 > 0 && nsample2 > 0 && B <= 65535, PEM_E_SHAPE);
    PEM_REQUIRE(nsample1 <= 128 && nsample2 <= 128 && N <= 65535, PEM_E_UNSUPPORTED);
    const bool in_smem = N <= 4096;  // whole cloud as float4: <= 64 KB
    const size_t n_pts = in_smem ? (size_t)(N + kPeTile - 1) / kPeTile * kPeTile : (size_t)kPeTile;
    const size_t smem = n_pts * sizeof(float4) +
                        (size_t)kPeCentres * (nsample1 + 2 + nsample2 + 2) * sizeof(unsigned short);
    const dim3 grid((N + kPeCentres - 1) / kPeCentres, B);
    if (in_smem) {
        static PemSmemCache cached(0);
        if (int rc = pem_ensure_dyn_smem(pe_group2_kernel<true, false>, smem, &cached)) return rc;
        pe_group2_kernel<true, false><<<grid, kPeThreads, smem, pem_cu_stream(stream)>>>(
            xyz, N, centre_eps, radius1 * radius1, nsample1, out1, radius2 * radius2, nsample2, out2,
            nullptr, 0, nullptr, nullptr);
    } else {
        static PemSmemCache cached(0);
        if (int rc = pem_ensure_dyn_smem(pe_group2_kernel<false, false>, smem, &cached)) return rc;
        pe_group2_kernel<false, false><<<grid, kPeThreads, smem, pem_cu_stream(stream)>>>(
            xyz, N, centre_eps, radius1 * radius1, nsample1, out1, radius2 * radius2, nsample2, out2,
            nullptr, 0, nullptr, nullptr);
    }
    return pem_launch_status();
}

namespace pem {
int pe_mask_tc_launch(const float *xyz, int B, int N, float eps, float thr, float rel_slack, int max_ctas,
                      uint32_t *masks, cudaStream_t st);  // pe_mask_tc.cu
}

extern "C" size_t pem_pe_masks_bytes(int B, int N) {
    if (B <= 0 || N <= 0) return 0;
    return (size_t)B * N * ((N + 31) / 32) * sizeof(uint32_t);
}

extern "C" int pem_pe_masks(const float *xyz, int B, int N, float centre_eps, float radius, float rel_margin,
                            int max_ctas, void *masks, size_t masks_bytes, pem_stream_t stream) {
    PEM_REQUIRE_PTR(xyz);
    PEM_REQUIRE_PTR(masks);
    PEM_REQUIRE(B > 0 && N > 0 && radius > 0.f && rel_margin >= 0.f && rel_margin <= 0.5f, PEM_E_SHAPE);
    PEM_REQUIRE(N <= 4096, PEM_E_UNSUPPORTED);
    PEM_REQUIRE(masks_bytes >= pem_pe_masks_bytes(B, N), PEM_E_WORKSPACE);
    return pe_mask_tc_launch(xyz, B, N, centre_eps, radius * radius, rel_margin, max_ctas, (uint32_t *)masks,
                             pem_cu_stream(stream));
}

static int pe_masked_launch(const float *xyz, int B, int N, float centre_eps, float radius1, int nsample1,
                            float *out1, int32_t *idx1, float radius2, int nsample2, float *out2, int32_t *idx2,
                            const void *masks, size_t masks_bytes, pem_stream_t stream) {
    PEM_REQUIRE_PTR(xyz);
    PEM_REQUIRE_PTR(masks);
    PEM_REQUIRE(B > 0 && N > 0 && nsample1 > 0 && nsample2 > 0 && B <= 65535, PEM_E_SHAPE);
    PEM_REQUIRE(nsample1 <= 128 && nsample2 <= 128 && N <= 4096, PEM_E_UNSUPPORTED);  // cloud in smem
    PEM_REQUIRE(masks_bytes >= pem_pe_masks_bytes(B, N), PEM_E_WORKSPACE);
    const int W = (N + 31) / 32;
    const size_t smem = (size_t)(N + 32 + 2) * sizeof(float4) +
                        (size_t)kPwWarps * (nsample1 + nsample2 + kPwCand) * sizeof(unsigned short);
    auto go = [&](auto kern, PemSmemCache *cache) {
        if (int rc = pem_ensure_dyn_smem(kern, smem, cache)) return rc;
        kern<<<dim3((N + kPwCentres - 1) / kPwCentres, B), 32 * kPwWarps, smem, pem_cu_stream(stream)>>>(
            xyz, N, centre_eps, radius1 * radius1, nsample1, out1, radius2 * radius2, nsample2, out2,
            (const uint32_t *)masks, W, idx1, idx2);
        return pem_launch_status();
    };
    static PemSmemCache c1(0), c2(0), c2pe(0), c4(0);
    if (W <= 32) return go(pe_group2_warp_kernel<1, 0, 0>, &c1);
    if (W <= 64 && nsample1 == 32 && nsample2 == 64) return go(pe_group2_warp_kernel<2, 32, 64>, &c2pe);  // the PEM shape
    if (W <= 64) return go(pe_group2_warp_kernel<2, 0, 0>, &c2);
    return go(pe_group2_warp_kernel<4, 0, 0>, &c4);
}

extern "C" int pem_pe_group2_masked(const float *xyz, int B, int N, float centre_eps, float radius1,
                                    int nsample1, float *out1, float radius2, int nsample2, float *out2,
                                    const void *masks, size_t masks_bytes, pem_stream_t stream) {
    PEM_REQUIRE_PTR(out1);
    PEM_REQUIRE_PTR(out2);
    return pe_masked_launch(xyz, B, N, centre_eps, radius1, nsample1, out1, nullptr, radius2, nsample2, out2,
                            nullptr, masks, masks_bytes, stream);
}

extern "C" int pem_pe_neighbours_masked(const float *xyz, int B, int N, float centre_eps, float radius1,
                                        int nsample1, int32_t *idx1, float radius2, int nsample2,
                                        int32_t *idx2, const void *masks, size_t masks_bytes,
                                        pem_stream_t stream) {
    PEM_REQUIRE_PTR(idx1);
    PEM_REQUIRE_PTR(idx2);
    return pe_masked_launch(xyz, B, N, centre_eps, radius1, nsample1, nullptr, idx1, radius2, nsample2, nullptr,
                            idx2, masks, masks_bytes, stream);
}

// 64-byte per-instance pose records (R 9 | t 3 | score | proposal id | instance id | pad): the
// unit the multi-GPU gather moves.
__global__ void pack_records_kernel(const float *__restrict__ R, const float *__restrict__ t,
                                    const float *__restrict__ score, const int32_t *__restrict__ pid,
                                    const float *__restrict__ inst, int B, float *__restrict__ rec) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 16) return;
    const int b = i >> 4, k = i & 15;
    float v = 0.f;
    if (k < 9) v = R[b * 9 + k];
    else if (k < 12) v = t[b * 3 + k - 9];
    else if (k == 12) v = score[b];
    else if (k == 13) v = pid ? (float)pid[b] : 0.f;
    else if (k == 14) v = inst ? inst[b] : (float)b;
    rec[i] = v;
}

extern "C" int pem_pack_records(const float *R, const float *t, const float *score,
                                const int32_t *proposal_id, const float *instance_id, int B,
                                float *records, pem_stream_t stream) {
    PEM_REQUIRE_PTR(R);
    PEM_REQUIRE_PTR(t);
    PEM_REQUIRE_PTR(score);
    PEM_REQUIRE_PTR(records);
    PEM_REQUIRE(B > 0, PEM_E_SHAPE);
    pack_records_kernel<<<(B * 16 + 255) / 256, 256, 0, pem_cu_stream(stream)>>>(R, t, score, proposal_id,
                                                                                instance_id, B, records);
    return pem_launch_status();
}
