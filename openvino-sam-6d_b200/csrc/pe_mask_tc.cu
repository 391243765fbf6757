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
// CTA = (cloud, block of <= 1024 columns j, a share of the 128-row tiles), warp-specialised like
// score_tc.cu (round 2; the round-1 form -- row threads that built the tile, loaded, waited and compared in
// lock step over two 256-column buffers -- took 5800 clocks per 128 x 1024 tile):
//   * the column block is the resident B operand (SWIZZLE_128B rows in shared memory);
//   * warps 0-1 issue the MMAs (M128 x N128 x K8, two per 128-column chunk; even chunks warp 0, odd
//     chunks warp 1) into four accumulator buffers (all 512 TMEM columns);
//   * warps 2-3 (64 threads, two tile rows each) write the next centre tile, double buffered.  The row's
//     limit rides in the two spare K slots, A = (-lim_hi, -lim_lo) against B = (1, 1), so the accumulator
//     holds D - lim and the candidate test is its SIGN BIT;
//   * warps 4-11 (two per TMEM lane quadrant; one takes the even chunks, the other the odd ones) load a chunk
//     in two 64-column halves, hand the buffer back when the second has landed and shift the sign bits into
//     four 32-bit mask words (one funnel shift per pair instead of compare + predicated OR), one 16-byte
//     store per row and chunk.
// When the cloud is a multiple of 1024 points every tile has eight chunks and the chunk loop is unrolled with
// constant barrier slots and parities (FIXED); other sizes take the generic loop.
#include "tc_ptx.cuh"
#include <type_traits>

namespace pem {
namespace tc {

constexpr int kMRows = 128;          // rows per tile = TMEM lanes
constexpr int kMChunk = 128;         // columns per accumulator buffer
constexpr int kMBufs = 4;            // accumulator buffers
constexpr int kMCols = 1024;         // columns per CTA (128 B per row of the B operand)
constexpr int kMFixedChunks = kMCols / kMChunk;
constexpr int kMIssuers = 2, kMBuilders = 64, kMReducers = 256;
constexpr int kMThreads = 32 * kMIssuers + kMBuilders + kMReducers;  // 12 warps
constexpr uint32_t kMIdesc = make_idesc_tf32(kMRows, kMChunk);

struct MaskSmem {
    uint8_t b[kMCols * 128];         // column block of the cloud
    uint8_t a[2][kMRows * 128];      // centre tiles
    float red[6][16];                // bounding-box reduction
    float box[6];
    uint64_t a_full[2], a_empty[2], tmem_full[kMBufs], tmem_empty[kMBufs];
    uint32_t tmem_base;
};

// A / B operand rows (see score_tc.cu): 14 of the 16 K slots carry the four hi/lo cross products
__device__ __forceinline__ void split3(float x, float y, float z, float (&h)[3], float (&l)[3]) {
    h[0] = to_tf32(x); h[1] = to_tf32(y); h[2] = to_tf32(z);
    l[0] = to_tf32(x - h[0]); l[1] = to_tf32(y - h[1]); l[2] = to_tf32(z - h[2]);
}

// 32 accumulator values -> one mask word: bit q = sign of r[q] (candidate when D - lim < 0)
__device__ __forceinline__ uint32_t sign_word(const uint32_t *r) {
    uint32_t m = 0u;
#pragma unroll
    for (int q = 31; q >= 0; --q) m = shift_in_sign(m, __uint_as_float(r[q]));
    return m;
}

template <bool FIXED>
__global__ void __launch_bounds__(kMThreads, 1)
pe_mask_tc_kernel(const float *__restrict__ xyz, int N, float eps, float thr, float rel_slack, int W,
                  int n_colblocks, int n_units, uint32_t *__restrict__ masks /* (B,N,W) */) {
    extern __shared__ uint8_t smem_raw[];
    MaskSmem &sm = *reinterpret_cast<MaskSmem *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = warp_id(), lane = lane_id(), tid = threadIdx.x;
    const bool is_mma = warp < kMIssuers, is_builder = warp >= kMIssuers && warp <= 3;
    const int n_tiles = (N + kMRows - 1) / kMRows;
    // work unit u = ((cloud * n_colblocks + column block) * n_tiles + row tile); a CTA owns a
    // contiguous range of units and rebuilds its B operand when the (cloud, column block) changes
    const int u_begin = (int)((long long)n_units * blockIdx.x / gridDim.x);
    const int u_end = (int)((long long)n_units * (blockIdx.x + 1) / gridDim.x);

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&sm.a_full[i]), kMBuilders);
            mbar_init(smem_u32(&sm.a_empty[i]), kMIssuers);
        }
        for (int i = 0; i < kMBufs; ++i) {
            mbar_init(smem_u32(&sm.tmem_full[i]), 1);
            mbar_init(smem_u32(&sm.tmem_empty[i]), kMReducers / 64);  // one reducer warp per lane quadrant
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
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
    const uint32_t full0 = smem_u32(&sm.tmem_full[0]), empty0 = smem_u32(&sm.tmem_empty[0]);
    uint32_t tile_seq = 0, chunk_seq = 0;  // barrier phases run on across segments
    int box_cloud = -1;

    for (int u = u_begin; u < u_end;) {
        const int bc = u / n_tiles, t0 = u - bc * n_tiles, t1 = min(n_tiles, t0 + (u_end - u));
        const int cb_idx = bc % n_colblocks, b = bc / n_colblocks;
        const int col0 = cb_idx * kMCols, ncols = min(kMCols, N - col0);
        const int n_chunks = FIXED ? kMFixedChunks : (ncols + kMChunk - 1) / kMChunk;
        const float *pb = xyz + (size_t)b * N * 3;
        u += t1 - t0;

        // ---- bounding box of the cloud (once per cloud): operands are taken relative to its centre
        //      (smaller magnitudes, smaller rounding), and its half diagonal bounds the filter's error
        if (box_cloud != b) {
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
            box_cloud = b;
        }
        __syncthreads();  // (also: every MMA of the previous segment has retired -- its accumulators were all
                          //  consumed -- before the B operand is rewritten below)
        const float ox = 0.5f * (sm.box[0] + sm.box[3]), oy = 0.5f * (sm.box[1] + sm.box[4]),
                    oz = 0.5f * (sm.box[2] + sm.box[5]);
        const float hx = sm.box[3] - ox, hy = sm.box[4] - oy, hz = sm.box[5] - oz;
        const float R2 = hx * hx + hy * hy + hz * hz + 3.f * eps * eps + 1e-30f;  // (half diagonal)^2
        // |computed - exact| <= ~32 * 2^-22 * R2 (14 split products, the split limit, fp32 accumulation);
        // ~3x headroom
        const float cut = thr * (1.0f + rel_slack) + 2.2e-5f * R2;

        // ---- B operand: the columns of this segment
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
                c3 = make_float4(sh, sl, 1.f, 1.f);   // K slots 14, 15: x (-lim_hi, -lim_lo) of the row
            } else {  // padding: never a candidate (D - lim = +1e30)
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
        const int nt = t1 - t0;

        if (is_mma) {
            // ================================ MMA issuers =================================
            const uint64_t b0 = umma_desc_sw128(smem_u32(sm.b));
            for (int tl = 0; tl < nt; ++tl, ++tile_seq) {
                const int buf = tile_seq & 1;
                mbar_wait(smem_u32(&sm.a_full[buf]), (tile_seq >> 1) & 1);
                tc_fence_after();
                const uint64_t ad = umma_desc_sw128(smem_u32(sm.a[buf]));
                if constexpr (FIXED) {
#pragma unroll
                    for (int c0_ = 0; c0_ < kMFixedChunks; c0_ += kMIssuers) {
                        const int c = c0_ + warp;
                        const uint32_t slot = 8u * (uint32_t)(c & 3), par = (uint32_t)(((c0_ >> 2) & 1) ^ 1);
                        mbar_wait(empty0 + slot, par);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t d = tmem_base + (uint32_t)((c & 3) * kMChunk);
                            const uint64_t bd = b0 + (uint64_t)(c * (kMChunk * 128 / 16));
                            tc_mma_tf32(d, ad, bd, kMIdesc, 0);          // K slots 0..7
                            tc_mma_tf32(d, ad + 2, bd + 2, kMIdesc, 1);  // K slots 8..15 (+32 bytes)
                            tc_commit(full0 + slot);
                            if (c0_ == kMFixedChunks - kMIssuers) tc_commit(smem_u32(&sm.a_empty[buf]));
                        }
                        __syncwarp();
                    }
                } else {
                    bool mine = false;
                    for (int c = 0; c < n_chunks; ++c, ++chunk_seq) {
                        if ((int)(chunk_seq & (kMIssuers - 1)) != warp) continue;
                        const bool last_of_mine = c + kMIssuers >= n_chunks;
                        mine = true;
                        const int cb = chunk_seq & (kMBufs - 1);
                        mbar_wait(smem_u32(&sm.tmem_empty[cb]), ((chunk_seq / kMBufs) & 1) ^ 1);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t d = tmem_base + (uint32_t)(cb * kMChunk);
                            const uint64_t bd = b0 + (uint64_t)(c * (kMChunk * 128 / 16));
                            tc_mma_tf32(d, ad, bd, kMIdesc, 0);
                            tc_mma_tf32(d, ad + 2, bd + 2, kMIdesc, 1);
                            tc_commit(smem_u32(&sm.tmem_full[cb]));
                            if (last_of_mine) tc_commit(smem_u32(&sm.a_empty[buf]));
                        }
                        __syncwarp();
                    }
                    if (!mine && lane == 0) mbar_arrive(smem_u32(&sm.a_empty[buf]));
                }
            }
        } else if (is_builder) {
            // ================================ centre-tile builders ========================
            const int e0 = tid - 32 * kMIssuers;
            for (int tl = 0; tl < nt; ++tl, ++tile_seq) {
                const int buf = tile_seq & 1;
                mbar_wait(smem_u32(&sm.a_empty[buf]), ((tile_seq >> 1) & 1) ^ 1);  // MMAs of the tile two back retired
#pragma unroll
                for (int r = 0; r < kMRows / kMBuilders; ++r) {
                    const int et = e0 + r * kMBuilders, i = (t0 + tl) * kMRows + et;
                    float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0, c2 = c0, c3 = c0;
                    if (i < N) {
                        const float *p = pb + 3 * (size_t)i;
                        const float qx = __fadd_rn(__ldg(p), eps) - ox, qy = __fadd_rn(__ldg(p + 1), eps) - oy,
                                    qz = __fadd_rn(__ldg(p + 2), eps) - oz;
                        const float lim = cut - (qx * qx + qy * qy + qz * qz);
                        const float lh = to_tf32(lim), ll = to_tf32(lim - lh);
                        float h[3], l[3];
                        split3(qx, qy, qz, h, l);
                        c0 = make_float4(h[0], h[1], h[2], h[0]);
                        c1 = make_float4(h[1], h[2], l[0], l[1]);
                        c2 = make_float4(l[2], l[0], l[1], l[2]);
                        c3 = make_float4(1.f, 1.f, -lh, -ll);
                    }
                    *reinterpret_cast<float4 *>(sm.a[buf] + sw128_offset(et, 0)) = c0;
                    *reinterpret_cast<float4 *>(sm.a[buf] + sw128_offset(et, 1)) = c1;
                    *reinterpret_cast<float4 *>(sm.a[buf] + sw128_offset(et, 2)) = c2;
                    *reinterpret_cast<float4 *>(sm.a[buf] + sw128_offset(et, 3)) = c3;
                }
                fence_proxy_async_smem();
                mbar_arrive(smem_u32(&sm.a_full[buf]));
            }
            chunk_seq += (uint32_t)(nt * n_chunks);
        } else {
            // ================================ reducers ====================================
            const int quad = warp & 3, part = (warp - 4) >> 2;
            const int et = quad * 32 + lane;  // tile row = TMEM lane
            const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
            const bool vec = (W & 3) == 0;
            auto take = [&](uint32_t full_bar, uint32_t parity, uint32_t empty_bar, uint32_t col, uint32_t *mrow,
                            int w0, bool live) {
                uint32_t lo[64], hi[64];
                mbar_wait(full_bar, parity);
                tc_fence_after();
                tmem_ld64_issue(lane_base + col, lo);
                tmem_ld64_wait(lo);
                tmem_ld64_issue(lane_base + col + 64u, hi);
                const uint32_t m0 = sign_word(lo), m1 = sign_word(lo + 32);
                tmem_ld64_wait(hi);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(empty_bar);
                const uint32_t m2 = sign_word(hi), m3 = sign_word(hi + 32);
                if (live) {
                    if (vec && w0 + 3 < W) {  // one 16-byte store (rows are 16-byte aligned)
                        *reinterpret_cast<uint4 *>(mrow + w0) = make_uint4(m0, m1, m2, m3);
                    } else {
                        if (w0 < W) mrow[w0] = m0;
                        if (w0 + 1 < W) mrow[w0 + 1] = m1;
                        if (w0 + 2 < W) mrow[w0 + 2] = m2;
                        if (w0 + 3 < W) mrow[w0 + 3] = m3;
                    }
                }
            };
            if constexpr (FIXED) {
                auto tiles = [&](auto part_tag) {
                    constexpr int PART = decltype(part_tag)::value;
                    for (int tl = 0; tl < nt; ++tl, ++tile_seq) {
                        const int i = (t0 + tl) * kMRows + et;
                        uint32_t *mrow = masks + ((size_t)b * N + min(i, N - 1)) * W;
#pragma unroll
                        for (int j = 0; j < kMFixedChunks / 2; ++j) {
                            const int c = 2 * j + PART;
                            const uint32_t slot = (uint32_t)(c & 3);
                            take(full0 + 8u * slot, (uint32_t)((j >> 1) & 1), empty0 + 8u * slot, slot * kMChunk, mrow,
                                 (col0 + c * kMChunk) >> 5, i < N);
                        }
                    }
                };
                if (part == 0) tiles(std::integral_constant<int, 0>{});
                else tiles(std::integral_constant<int, 1>{});
            } else {
                for (int tl = 0; tl < nt; ++tl, ++tile_seq) {
                    const int i = (t0 + tl) * kMRows + et;
                    uint32_t *mrow = masks + ((size_t)b * N + min(i, N - 1)) * W;
                    for (int c = 0; c < n_chunks; ++c, ++chunk_seq) {
                        if ((int)(chunk_seq & 1) != part) continue;
                        const uint32_t slot = chunk_seq & (kMBufs - 1);
                        take(full0 + 8u * slot, (chunk_seq / kMBufs) & 1, empty0 + 8u * slot, slot * kMChunk, mrow,
                             (col0 + c * kMChunk) >> 5, i < N);
                    }
                }
            }
        }
        // (the next segment's first barrier keeps B and the box from being rebuilt under a reader)
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
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
    if (N % kMCols == 0) {  // every tile of every segment has eight chunks: the unrolled form
        static PemSmemCache cached(0);
        if (int rc = pem_ensure_dyn_smem(pe_mask_tc_kernel<true>, smem, &cached)) return rc;
        pe_mask_tc_kernel<true><<<grid, kMThreads, smem, st>>>(xyz, N, eps, thr, rel_slack, W, n_colblocks, n_units,
                                                               masks);
    } else {
        static PemSmemCache cached(0);
        if (int rc = pem_ensure_dyn_smem(pe_mask_tc_kernel<false>, smem, &cached)) return rc;
        pe_mask_tc_kernel<false><<<grid, kMThreads, smem, st>>>(xyz, N, eps, thr, rel_slack, W, n_colblocks, n_units,
                                                                masks);
    }
    return pem_launch_status();
}

}  // namespace pem
