// Ball-query candidate masks on the tensor cores.
//
// The PositionalEncoding grouping (fine_point_matching.py:113-137) ball-queries every point of a
// 2048-point cloud against the whole cloud: 4.2 M squared distances per cloud, which the exact
// SIMT scan (pointnet2_ops.cu) pays for with ~13 instructions each because the reference's
// arithmetic (separately rounded dx^2 + dy^2 + dz^2) must be reproduced bit for bit.  Only ~4 % of
// the pairs are anywhere near the larger ball, so this kernel runs a cheap, slightly GENEROUS
// filter first:
//
//   D[i, j] = |p_j|^2 - 2 c_i . p_j      (one row of a GEMM, tcgen05.mma kind::tf32, split-TF32
//                                          operands as in score_tc.cu: ~2^-22 relative)
//   bit j of mask[i]  =  D[i, j] < thr + slack - |c_i|^2
//
// with slack covering the filter's own rounding (a bound from the cloud's extent) plus, when the
// cloud is going to be moved rigidly before it is grouped, the rounding of that motion.  The exact
// kernel then re-tests only the set bits on the (moved) cloud, so its results are unchanged.
//
// CTA = (cloud, block of <= 1024 columns j, a share of the 128-row tiles): the column block is the
// resident B operand (SWIZZLE_128B rows in shared memory), row tiles are written by the row
// threads themselves, accumulator chunks (128 x 256) are double buffered in TMEM, and each row
// thread turns its lane of a chunk into four 32-bit mask words.
#include "tc_ptx.cuh"

namespace pem {
namespace tc {

constexpr int kMRows = 128;          // rows per tile = TMEM lanes
constexpr int kMChunk = 256;         // columns per accumulator buffer
constexpr int kMCols = 1024;         // columns per CTA (128 B per row of the B operand)
constexpr int kMRowThreads = 256;
constexpr int kMThreads = 64 + kMRowThreads;
constexpr uint32_t kMIdesc = make_idesc_tf32(kMRows, kMChunk);

struct MaskSmem {
    uint8_t b[kMCols * 128];         // column block of the cloud
    uint8_t a[2][kMRows * 128];      // centre tiles
    float red[6][16];                // bounding-box reduction
    float box[6];
    uint64_t a_full[2], a_empty[2], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

// A / B operand rows (see score_tc.cu): 14 of the 16 K slots carry the four hi/lo cross products
__device__ __forceinline__ void split3(float x, float y, float z, float (&h)[3], float (&l)[3]) {
    h[0] = to_tf32(x); h[1] = to_tf32(y); h[2] = to_tf32(z);
    l[0] = to_tf32(x - h[0]); l[1] = to_tf32(y - h[1]); l[2] = to_tf32(z - h[2]);
}

__global__ void __launch_bounds__(kMThreads, 1)
pe_mask_tc_kernel(const float *__restrict__ xyz, int N, float eps, float thr, float rel_slack, int W,
                  int n_colblocks, int n_units, uint32_t *__restrict__ masks /* (B,N,W) */) {
    extern __shared__ uint8_t smem_raw[];
    MaskSmem &sm = *reinterpret_cast<MaskSmem *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = warp_id(), lane = lane_id(), tid = threadIdx.x;
    const int n_tiles = (N + kMRows - 1) / kMRows;
    // work unit u = ((cloud * n_colblocks + column block) * n_tiles + row tile); a CTA owns a
    // contiguous range of units and rebuilds its B operand when the (cloud, column block) changes
    const int u_begin = (int)((long long)n_units * blockIdx.x / gridDim.x);
    const int u_end = (int)((long long)n_units * (blockIdx.x + 1) / gridDim.x);

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&sm.a_full[i]), kMRowThreads);
            mbar_init(smem_u32(&sm.a_empty[i]), 1);
            mbar_init(smem_u32(&sm.tmem_full[i]), 1);
            mbar_init(smem_u32(&sm.tmem_empty[i]), kMRowThreads / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&sm.tmem_base)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;
    uint32_t tile_seq = 0, chunk_seq = 0;  // barrier phases run on across segments

    for (int u = u_begin; u < u_end;) {
        const int bc = u / n_tiles, t0 = u - bc * n_tiles, t1 = min(n_tiles, t0 + (u_end - u));
        const int cb_idx = bc % n_colblocks, b = bc / n_colblocks;
        const int col0 = cb_idx * kMCols, ncols = min(kMCols, N - col0);
        const int n_chunks = (ncols + kMChunk - 1) / kMChunk;
        const float *pb = xyz + (size_t)b * N * 3;
        u += t1 - t0;

        // ---- bounding box of the cloud: operands are taken relative to its centre (smaller
        //      magnitudes, smaller rounding), and its half diagonal bounds the filter's error
        {
            float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (int k = tid; k < N; k += kMThreads)
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const float v = __ldg(pb + 3 * k + a);
                    lo[a] = fminf(lo[a], v);
                    hi[a] = fmaxf(hi[a], v);
                }
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                lo[a] = warp_min(lo[a]);
                hi[a] = warp_max(hi[a]);
                if (lane == 0) { sm.red[a][warp] = lo[a]; sm.red[3 + a][warp] = hi[a]; }
            }
            __syncthreads();
            if (tid < 6) {
                float v = sm.red[tid][0];
                for (int w = 1; w < kMThreads / 32; ++w)
                    v = tid < 3 ? fminf(v, sm.red[tid][w]) : fmaxf(v, sm.red[tid][w]);
                sm.box[tid] = v;
            }
            __syncthreads();
        }
        const float ox = 0.5f * (sm.box[0] + sm.box[3]), oy = 0.5f * (sm.box[1] + sm.box[4]),
                    oz = 0.5f * (sm.box[2] + sm.box[5]);
        const float hx = sm.box[3] - ox, hy = sm.box[4] - oy, hz = sm.box[5] - oz;
        const float R2 = hx * hx + hy * hy + hz * hz + 3.f * eps * eps + 1e-30f;  // (half diagonal)^2
        // |computed - exact| <= ~30 * 2^-22 * R2 (14 split products + fp32 accumulation); 3x headroom
        const float cut = thr * (1.0f + rel_slack) + 2.2e-5f * R2;

        // ---- B operand: the columns of this segment (every MMA of the previous one has retired:
        //      its row threads consumed all of its accumulators before the barrier below)
        for (int m = tid; m < n_chunks * kMChunk; m += kMThreads) {
            float4 c0, c1, c2, c3;
            if (m < ncols) {
                const float *p = pb + 3 * (size_t)(col0 + m);
                const float x = __ldg(p) - ox, y = __ldg(p + 1) - oy, z = __ldg(p + 2) - oz;
                float h[3], l[3];
                split3(x, y, z, h, l);
                const float s2 = x * x + y * y + z * z, sh = to_tf32(s2), sl = to_tf32(s2 - sh);
                c0 = make_float4(-2.f * h[0], -2.f * h[1], -2.f * h[2], -2.f * l[0]);
                c1 = make_float4(-2.f * l[1], -2.f * l[2], -2.f * h[0], -2.f * h[1]);
                c2 = make_float4(-2.f * h[2], -2.f * l[0], -2.f * l[1], -2.f * l[2]);
                c3 = make_float4(sh, sl, 0.f, 0.f);
            } else {  // padding: never a candidate
                c0 = c1 = c2 = make_float4(0.f, 0.f, 0.f, 0.f);
                c3 = make_float4(1e30f, 0.f, 0.f, 0.f);
            }
            *reinterpret_cast<float4 *>(sm.b + sw128_offset(m, 0)) = c0;
            *reinterpret_cast<float4 *>(sm.b + sw128_offset(m, 1)) = c1;
            *reinterpret_cast<float4 *>(sm.b + sw128_offset(m, 2)) = c2;
            *reinterpret_cast<float4 *>(sm.b + sw128_offset(m, 3)) = c3;
        }
        fence_proxy_async_smem();
        __syncthreads();

        if (warp == 1) {
            // ================================ MMA issuer ==================================
            const uint64_t b0 = umma_desc_sw128(smem_u32(sm.b));
            for (int tile = t0; tile < t1; ++tile, ++tile_seq) {
                const int buf = tile_seq & 1;
                mbar_wait(smem_u32(&sm.a_full[buf]), (tile_seq >> 1) & 1);
                tc_fence_after();
                const uint64_t ad = umma_desc_sw128(smem_u32(sm.a[buf]));
                for (int c = 0; c < n_chunks; ++c, ++chunk_seq) {
                    const int cb = chunk_seq & 1;
                    mbar_wait(smem_u32(&sm.tmem_empty[cb]), ((chunk_seq >> 1) & 1) ^ 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t d = tmem_base + (uint32_t)(cb * kMChunk);
                        const uint64_t bd = b0 + (uint64_t)(c * (kMChunk * 128 / 16));
                        tc_mma_tf32(d, ad, bd, kMIdesc, 0);          // K slots 0..7
                        tc_mma_tf32(d, ad + 2, bd + 2, kMIdesc, 1);  // K slots 8..15 (+32 bytes)
                        tc_commit(smem_u32(&sm.tmem_full[cb]));
                        if (c == n_chunks - 1) tc_commit(smem_u32(&sm.a_empty[buf]));
                    }
                    __syncwarp();
                }
            }
        } else if (warp >= 2) {
            // ================================ row threads =================================
            const int quad = warp & 3, half = (warp - 2) >> 2;
            const int et = quad * 32 + lane;  // tile row = TMEM lane

            // centre of row `et` of a tile -> this thread's half of the A row; returns cut - |c|^2
            auto build = [&](int tile, uint32_t seq) {
                const int buf = seq & 1;
                mbar_wait(smem_u32(&sm.a_empty[buf]), ((seq >> 1) & 1) ^ 1);
                const int i = tile * kMRows + et;
                float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
                float lim = -INFINITY;
                if (i < N) {
                    const float *p = pb + 3 * (size_t)i;
                    const float qx = __fadd_rn(__ldg(p), eps) - ox, qy = __fadd_rn(__ldg(p + 1), eps) - oy,
                                qz = __fadd_rn(__ldg(p + 2), eps) - oz;
                    lim = cut - (qx * qx + qy * qy + qz * qz);
                    float h[3], l[3];
                    split3(qx, qy, qz, h, l);
                    if (half == 0) {
                        c0 = make_float4(h[0], h[1], h[2], h[0]);
                        c1 = make_float4(h[1], h[2], l[0], l[1]);
                    } else {
                        c0 = make_float4(l[2], l[0], l[1], l[2]);
                        c1 = make_float4(1.f, 1.f, 0.f, 0.f);
                    }
                }
                *reinterpret_cast<float4 *>(sm.a[buf] + sw128_offset(et, 2 * half)) = c0;
                *reinterpret_cast<float4 *>(sm.a[buf] + sw128_offset(et, 2 * half + 1)) = c1;
                fence_proxy_async_smem();
                mbar_arrive(smem_u32(&sm.a_full[buf]));
                return lim;
            };

            float lim_next = build(t0, tile_seq);
            for (int tile = t0; tile < t1; ++tile, ++tile_seq) {
                const float lim = lim_next;
                if (tile + 1 < t1) lim_next = build(tile + 1, tile_seq + 1);
                const int i = tile * kMRows + et;
                uint32_t *mrow = masks + ((size_t)b * N + min(i, N - 1)) * W;
                for (int c = 0; c < n_chunks; ++c, ++chunk_seq) {
                    const int cb = chunk_seq & 1;
                    mbar_wait(smem_u32(&sm.tmem_full[cb]), (chunk_seq >> 1) & 1);
                    tc_fence_after();
                    uint32_t wds[4] = {0u, 0u, 0u, 0u};
                    {
                        float v[kMChunk / 2];
                        tmem_ld128(tmem_base + ((uint32_t)(quad * 32) << 16) +
                                       (uint32_t)(cb * kMChunk + half * (kMChunk / 2)), v);
#pragma unroll
                        for (int q = 0; q < kMChunk / 2; ++q)
                            if (v[q] < lim) wds[q >> 5] |= 1u << (q & 31);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&sm.tmem_empty[cb]));
                    if (i < N) {
                        const int w0 = (col0 + c * kMChunk + half * (kMChunk / 2)) >> 5;  // first of 4 words
                        if ((W & 3) == 0 && w0 + 3 < W) {  // one 16-byte store (rows are 16-byte aligned)
                            *reinterpret_cast<uint4 *>(mrow + w0) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (w0 + k < W) mrow[w0 + k] = wds[k];
                        }
                    }
                }
            }
        }
        // keep the three roles in step: the issuer counts the same tiles and chunks
        if (warp == 0) {
            const uint32_t nt = (uint32_t)(t1 - t0);
            tile_seq += nt;
            chunk_seq += nt * (uint32_t)n_chunks;
        }
        __syncthreads();  // B and the box are rebuilt by the next segment
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u)
                     : "memory");
    }
}

}  // namespace tc

// masks (B,N,W) with W = ceil(N/32): bit j of row i set when point j MAY lie within sqrt(thr) of
// centre i (never a false negative for distances changed by less than rel_slack * thr afterwards)
int pe_mask_tc_launch(const float *xyz, int B, int N, float eps, float thr, float rel_slack, int max_ctas,
                      uint32_t *masks, cudaStream_t st) {
    using namespace tc;
    const int sms = pem_num_sms();
    const int W = (N + 31) / 32;
    const int n_colblocks = (N + kMCols - 1) / kMCols, n_tiles = (N + kMRows - 1) / kMRows;
    const int n_units = B * n_colblocks * n_tiles;
    int grid = min(n_units, sms);  // one CTA per SM (all of TMEM, 160 KB of shared memory)
    if (max_ctas > 0) grid = min(grid, max_ctas);
    const size_t smem = sizeof(MaskSmem) + 1024;
    static PemSmemCache cached(0);
    if (int rc = pem_ensure_dyn_smem(pe_mask_tc_kernel, smem, &cached)) return rc;
    pe_mask_tc_kernel<<<grid, kMThreads, smem, st>>>(xyz, N, eps, thr, rel_slack, W, n_colblocks, n_units, masks);
    return pem_launch_status();
}

}  // namespace pem
