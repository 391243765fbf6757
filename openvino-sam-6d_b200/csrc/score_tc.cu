// Hypothesis scoring (MU:267-271) on the tensor cores.
//
//   score_h = nfg / (sum_i min_m || (p_i - t_h) R_h - m || + 1e-8)      i over the foreground points
//
// is a nearest-neighbour search of ~150 transformed points against ~1024 model points for each of
// B x 300 hypotheses: 7.5e8 point pairs per step.  The squared distance expands to
// |q|^2 + (|m|^2 - 2 q.m), and the bracket is one row of a GEMM with K = 3 (+1 for |m|^2):
//
//   D[row, m] = sum_k A[row, k] * Bm[m, k]        row = (hypothesis, foreground point)
//
// computed with tcgen05.mma kind::tf32.  fp32-class accuracy comes from splitting every coordinate
// into two TF32 values (hi + lo, 22 significant bits) and giving each of the four cross products its
// own K slots -- 14 of the 16 K values of two K = 8 instructions:
//
//   A  row: qh.x qh.y qh.z | qh.x qh.y qh.z | ql.x ql.y ql.z | ql.x ql.y ql.z | 1    1    0 0
//   Bm row: -2mh.x ..      | -2ml.x ..      | -2mh.x ..      | -2ml.x ..      | s_hi s_lo 0 0     s = |m|^2
//
// One persistent CTA per (instance, slice of its hypotheses), warp-specialised:
//   * the model operand (1024 rows x 64 used bytes, SWIZZLE_128B K-major rows) is built once in
//     shared memory; rows of all hypotheses of the slice are packed back to back into 128-row
//     tiles, so no tensor work is wasted on padding 151 points to 256;
//   * warps 0-1 issue M128 x N128 x K8 MMAs, 2 per 128-model-point chunk (even chunks warp 0, odd chunks
//     warp 1: one warp walking its ~40 dependent instructions per chunk could not keep the reducers fed), into
//     FOUR accumulator buffers (all 512 TMEM columns), i.e. up to three chunks ahead of the reduction;
//   * warps 2-3 ("builders", 64 threads for the 128 tile rows) transform their points and write the A rows
//     (generic stores + proxy fence) of the next tile, double buffered, and leave |q|^2 in a small ring;
//   * warps 4-11 ("reducers", two per TMEM lane quadrant: one takes the even chunks, the other the odd
//     ones, all 128 columns) load their lane of an accumulator chunk, hand the buffer back the moment the
//     load has landed and fold it with 3-input minima (8 independent chains); the two warps of a scheduler
//     wait on different barriers, so one folds while the other's load is in flight;
//   * per-warp row minima (+ |q|^2) land in shared memory, and each hypothesis is summed in a fixed order.
// What bounds it (tools/tmem_ld_rate.cu, tools/minmax_rate.cu on this GPU): tcgen05.ld 32x32b delivers
// ~950 B/clk/SM with two warps per quadrant, so the accumulator read is NOT the limit (the round-1 form
// of this kernel -- load, wait, fold in lock step, two buffers -- sat at 128 B/clk/SM and was misread as
// port bound); FMNMX3 issues at 0.4 per clock per scheduler = 102 folded values per clock per SM, i.e.
// ~1290 clocks per 128 x 1024 tile, next to 1024 clocks of MMAs.  What the round-1 form and a first
// warp-specialised cut really spent their time on was the hand-shake itself: ~50 dependent integer
// instructions per chunk (barrier slot and parity from running counters) at ~6 clocks each with two warps
// per scheduler (ncu source page; with MMAs, loads and minima all switched off the kernel still took 3/4
// of its time).  Hence the FIXED form: when every model block is exactly eight chunks (Nm = 1024, 2048,
// ...) a tile's chunk loop is unrolled with every barrier slot, parity and TMEM column a compile-time
// constant (8 uses of 4 buffers per tile = an even number of phases per buffer and tile).  Other shapes take the
// generic loop.
// More than 1024 model points are taken in blocks of 1024: the tiles of a batch of hypotheses run
// once per block and the running minima wait in shared memory.  The SIMT kernel (coarse.cu) remains
// for shapes this one does not take (more than 4096 model points or 256 query points).
#include "tc_ptx.cuh"
#include <type_traits>

namespace pem {
namespace tc {

constexpr int kSRows = 128;          // rows per tile = TMEM lanes
constexpr int kSChunk = 128;         // model points per accumulator buffer
constexpr int kSBufs = 4;            // accumulator buffers (4 x 128 = all 512 TMEM columns)
constexpr int kSMaxNm = 1024;        // model points held in shared memory (128 B per row)
constexpr int kSFixedChunks = kSMaxNm / kSChunk;  // chunks per tile of the unrolled form
constexpr int kSMaxN1 = 256;         // query points per instance
constexpr int kSBatchHyps = 40;      // hypotheses per batch (their R | t staged in shared memory)
constexpr int kSRowCap = 6144;       // rows (hypothesis, point) of a batch whose minima fit in shared memory
constexpr int kSIssuers = 2;        // MMA-issuing warps: warp w takes the chunks whose running number has parity w
constexpr int kSBuilders = 64, kSReducers = 256, kSThreads = 32 * kSIssuers + kSBuilders + kSReducers;  // 12 warps: 168 registers
constexpr uint32_t kSIdesc = make_idesc_tf32(kSRows, kSChunk);

struct ScoreSmem {
    uint8_t b[kSMaxNm * 128];        // model operand
    uint8_t a[2][kSRows * 128];      // query tiles
    float dmin[2][kSRowCap];         // per column half: running minimum (+ |q|^2 after the last model block)
    float q2[8][kSRows];             // |q|^2 of the rows of the tiles in flight
    float px[kSMaxN1], py[kSMaxN1], pz[kSMaxN1];  // compacted foreground points
    float Rt[kSBatchHyps][12];       // R (row-major) | t of the hypotheses of the batch
    int nfg;
    uint64_t a_full[2], a_empty[2], tmem_full[kSBufs], tmem_empty[kSBufs];
    uint32_t tmem_base;
};

__device__ __forceinline__ void builder_barrier() { asm volatile("bar.sync 2, %0;" ::"n"(kSBuilders) : "memory"); }
__device__ __forceinline__ void reducer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kSReducers) : "memory"); }

// 64 accumulator values into 8 running minima (8 independent chains of 3-input minima)
__device__ __forceinline__ void fold64(const uint32_t (&r)[64], float (&m)[8]) {
#pragma unroll
    for (int q = 0; q < 64; q += 16) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            m[k] = fminf(fminf(m[k], __uint_as_float(r[q + 2 * k])), __uint_as_float(r[q + 2 * k + 1]));
    }
}
__device__ __forceinline__ float min8(const float (&m)[8]) {
    return fminf(fminf(fminf(m[0], m[1]), fminf(m[2], m[3])), fminf(fminf(m[4], m[5]), fminf(m[6], m[7])));
}

// FIXED: every model block has exactly kSFixedChunks chunks (Nm a multiple of 1024 after padding the last
// chunk: 896 < Nm - 1024 k <= 1024 for every block), so chunk_seq == 8 * tile_seq throughout and the
// barrier slot (c & 3) and parity (c >> 2) of chunk c of a tile are compile-time constants.
template <bool FIXED>
__global__ void __launch_bounds__(kSThreads, 1)
score_tc_kernel(const float *__restrict__ R, const float *__restrict__ t, const int32_t *__restrict__ sel,
                int P, int H, const float *__restrict__ pts1, const float *__restrict__ w1, int N1,
                const float *__restrict__ model, int Nm, int cpi, float *__restrict__ scores) {
    extern __shared__ uint8_t smem_raw[];
    ScoreSmem &sm = *reinterpret_cast<ScoreSmem *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = warp_id(), lane = lane_id(), tid = threadIdx.x;
    const int b = blockIdx.x / cpi, slice = blockIdx.x - b * cpi;
    const int hpc = (H + cpi - 1) / cpi;                       // hypotheses per CTA
    const int h_begin = slice * hpc, h_end = min(H, h_begin + hpc);
    const bool is_mma = warp < kSIssuers, is_builder = warp >= kSIssuers && warp <= 3, is_reducer = warp >= 4;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&sm.a_full[i]), kSBuilders);
            mbar_init(smem_u32(&sm.a_empty[i]), kSIssuers);
        }
        for (int i = 0; i < kSBufs; ++i) {
            mbar_init(smem_u32(&sm.tmem_full[i]), 1);
            mbar_init(smem_u32(&sm.tmem_empty[i]), kSReducers / 64);   // one reducer warp per lane quadrant
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
    const int n_blocks = (Nm + kSMaxNm - 1) / kSMaxNm;  // model blocks of <= 1024 points
    // the model operand: block `mblk` of the model points as B rows (all threads)
    auto build_model = [&](int mblk, int first_thread) {
        const float *mb = model + (size_t)b * Nm * 3;
        const int m_lo = mblk * kSMaxNm, cnt = min(kSMaxNm, Nm - m_lo);
        const int padded = (cnt + kSChunk - 1) / kSChunk * kSChunk;
        for (int m = tid - first_thread; m < padded; m += kSThreads - first_thread) {
            float4 c0, c1, c2, c3;
            if (m < cnt) {
                const float *p = mb + 3 * (size_t)(m_lo + m);
                const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
                const float xh = to_tf32(x), yh = to_tf32(y), zh = to_tf32(z);
                const float xl = to_tf32(x - xh), yl = to_tf32(y - yh), zl = to_tf32(z - zh);
                const float s = x * x + y * y + z * z, sh = to_tf32(s), sl = to_tf32(s - sh);
                c0 = make_float4(-2.f * xh, -2.f * yh, -2.f * zh, -2.f * xl);
                c1 = make_float4(-2.f * yl, -2.f * zl, -2.f * xh, -2.f * yh);
                c2 = make_float4(-2.f * zh, -2.f * xl, -2.f * yl, -2.f * zl);
                c3 = make_float4(sh, sl, 0.f, 0.f);
            } else {  // padding up to a whole chunk: a point that never wins the minimum
                c0 = c1 = c2 = make_float4(0.f, 0.f, 0.f, 0.f);
                c3 = make_float4(1e30f, 0.f, 0.f, 0.f);
            }
            *reinterpret_cast<float4 *>(sm.b + sw128_offset(m, 0)) = c0;
            *reinterpret_cast<float4 *>(sm.b + sw128_offset(m, 1)) = c1;
            *reinterpret_cast<float4 *>(sm.b + sw128_offset(m, 2)) = c2;
            *reinterpret_cast<float4 *>(sm.b + sw128_offset(m, 3)) = c3;
        }
        fence_proxy_async_smem();
    };

    // ---- per-CTA setup: foreground points (ordered compaction)
    // (warp 0 compacts -- seven dependent rounds of global loads -- while the other warps build the first model block)
    if (warp != 0) build_model(0, 32);
    if (warp == 0) {
        const float *wb = w1 + (size_t)b * N1, *pb = pts1 + (size_t)b * N1 * 3;
        int n = 0;
        for (int base = 0; base < N1; base += 32) {
            const int i = base + lane;
            const bool fg = i < N1 && __ldg(wb + i) != 0.f;
            const unsigned bal = __ballot_sync(kFull, fg);
            if (fg) {
                const int o = n + __popc(bal & ((1u << lane) - 1u));
                sm.px[o] = __ldg(pb + 3 * i);
                sm.py[o] = __ldg(pb + 3 * i + 1);
                sm.pz[o] = __ldg(pb + 3 * i + 2);
            }
            n += __popc(bal);
        }
        if (lane == 0) sm.nfg = n;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;
    const int nfg = sm.nfg;

    if (nfg > 0 && h_begin < h_end) {
        // Every warp walks the same (batch of hypotheses, model block) sequence; with more than 1024
        // model points the tiles of a batch are run once per model block and the running minimum of
        // each row waits in `dmin` in between.  Barrier phases (tile_seq, chunk_seq) run on throughout.
        uint32_t tile_seq = 0, chunk_seq = 0;
        int loaded = 0;  // model block currently in shared memory (block 0 was built during the setup)
        const int batch = max(1, min(kSBatchHyps, kSRowCap / nfg));
        const uint32_t full0 = smem_u32(&sm.tmem_full[0]), empty0 = smem_u32(&sm.tmem_empty[0]);
        for (int hb = h_begin; hb < h_end; hb += batch) {
            const int nb = min(h_end, hb + batch) - hb, rows = nb * nfg;
            const int n_tiles = (rows + kSRows - 1) / kSRows;
            if (is_builder) {
                builder_barrier();  // every builder is done with the previous batch's R | t
                for (int k = tid - 32 * kSIssuers; k < nb * 12; k += kSBuilders) {  // stage R | t of the batch
                    const int hh = k / 12, e = k - hh * 12, h = hb + hh;
                    const size_t src = (size_t)b * P + (sel ? __ldg(sel + (size_t)b * H + h) : h);
                    sm.Rt[hh][e] = e < 9 ? __ldg(R + src * 9 + e) : __ldg(t + src * 3 + (e - 9));
                }
                builder_barrier();
            }
            for (int mblk = 0; mblk < n_blocks; ++mblk) {
                const int n_chunks = FIXED ? kSFixedChunks
                                           : (min(kSMaxNm, Nm - mblk * kSMaxNm) + kSChunk - 1) / kSChunk;
                if (loaded != mblk) {  // (every MMA that read the previous block has retired: its
                    __syncthreads();   //  accumulators were all consumed before this barrier)
                    build_model(mblk, 0);
                    loaded = mblk;
                    __syncthreads();
                }
                const bool first = mblk == 0, last = mblk == n_blocks - 1;

                if (is_mma) {
                    // ================================ MMA issuer ==================================
                    const uint64_t b0 = umma_desc_sw128(smem_u32(sm.b));
                    for (int tl = 0; tl < n_tiles; ++tl, ++tile_seq) {
                        const int buf = tile_seq & 1;
                        mbar_wait(smem_u32(&sm.a_full[buf]), (tile_seq >> 1) & 1);
                        tc_fence_after();
                        const uint64_t ad = umma_desc_sw128(smem_u32(sm.a[buf]));
                        if constexpr (FIXED) {
#pragma unroll
                            for (int c0 = 0; c0 < kSFixedChunks; c0 += kSIssuers) {
                                const int c = c0 + warp;  // (kSFixedChunks is even: the parity of c is the warp's)
                                const uint32_t slot = 8u * (uint32_t)(c & 3), par = (uint32_t)(((c0 >> 2) & 1) ^ 1);
                                mbar_wait(empty0 + slot, par);
                                tc_fence_after();
                                if (elect_one()) {
                                    const uint32_t d = tmem_base + (uint32_t)((c & 3) * kSChunk);
                                    const uint64_t bd = b0 + (uint64_t)(c * (kSChunk * 128 / 16));
                                    tc_mma_tf32(d, ad, bd, kSIdesc, 0);          // K slots 0..7
                                    tc_mma_tf32(d, ad + 2, bd + 2, kSIdesc, 1);  // K slots 8..15 (+32 bytes)
                                    tc_commit(full0 + slot);
                                    if (c0 == kSFixedChunks - kSIssuers) tc_commit(smem_u32(&sm.a_empty[buf]));
                                }
                                __syncwarp();
                            }
                        } else {
                            bool mine = false;
                            for (int c = 0; c < n_chunks; ++c, ++chunk_seq) {
                                if ((int)(chunk_seq & (kSIssuers - 1)) != warp) continue;
                                const bool last_of_mine = c + kSIssuers >= n_chunks;
                                mine = true;
                                const int cb = chunk_seq & (kSBufs - 1);
                                mbar_wait(smem_u32(&sm.tmem_empty[cb]), ((chunk_seq / kSBufs) & 1) ^ 1);
                                tc_fence_after();
                                if (elect_one()) {
                                    const uint32_t d = tmem_base + (uint32_t)(cb * kSChunk);
                                    const uint64_t bd = b0 + (uint64_t)(c * (kSChunk * 128 / 16));
                                    tc_mma_tf32(d, ad, bd, kSIdesc, 0);          // K slots 0..7
                                    tc_mma_tf32(d, ad + 2, bd + 2, kSIdesc, 1);  // K slots 8..15 (+32 bytes)
                                    tc_commit(smem_u32(&sm.tmem_full[cb]));
                                    if (last_of_mine) tc_commit(smem_u32(&sm.a_empty[buf]));
                                }
                                __syncwarp();
                            }
                            // a warp without a chunk in this tile still owes the tile's a_empty arrival
                            if (!mine && lane == 0) mbar_arrive(smem_u32(&sm.a_empty[buf]));
                        }
                    }
                } else if (is_builder) {
                    // ================================ A-tile builders ==============================
                    for (int tl = 0; tl < n_tiles; ++tl, ++tile_seq) {
                        const int buf = tile_seq & 1;
                        mbar_wait(smem_u32(&sm.a_empty[buf]), ((tile_seq >> 1) & 1) ^ 1);  // MMAs of the tile two back retired
                        for (int et = tid - 32 * kSIssuers; et < kSRows; et += kSBuilders) {  // tile rows of this thread
                            const int row = tl * kSRows + et;
                            float4 c0, c1, c2, c3;
                            c0 = c1 = c2 = c3 = make_float4(0.f, 0.f, 0.f, 0.f);
                            float q2 = 0.f;
                            if (row < rows) {
                                const int hh = row / nfg, i = row - hh * nfg;
                                const float *Rm = sm.Rt[hh];
                                const float v0 = sm.px[i] - Rm[9], v1 = sm.py[i] - Rm[10], v2 = sm.pz[i] - Rm[11];
                                // (p1 - t) @ R : row vector times R (MU:267)
                                const float qx = v0 * Rm[0] + v1 * Rm[3] + v2 * Rm[6];
                                const float qy = v0 * Rm[1] + v1 * Rm[4] + v2 * Rm[7];
                                const float qz = v0 * Rm[2] + v1 * Rm[5] + v2 * Rm[8];
                                q2 = qx * qx + qy * qy + qz * qz;
                                const float xh = to_tf32(qx), yh = to_tf32(qy), zh = to_tf32(qz);
                                const float xl = to_tf32(qx - xh), yl = to_tf32(qy - yh), zl = to_tf32(qz - zh);
                                c0 = make_float4(xh, yh, zh, xh);
                                c1 = make_float4(yh, zh, xl, yl);
                                c2 = make_float4(zl, xl, yl, zl);
                                c3 = make_float4(1.f, 1.f, 0.f, 0.f);
                            }
                            *reinterpret_cast<float4 *>(sm.a[buf] + sw128_offset(et, 0)) = c0;
                            *reinterpret_cast<float4 *>(sm.a[buf] + sw128_offset(et, 1)) = c1;
                            *reinterpret_cast<float4 *>(sm.a[buf] + sw128_offset(et, 2)) = c2;
                            *reinterpret_cast<float4 *>(sm.a[buf] + sw128_offset(et, 3)) = c3;
                            // (the reducers read this slot when they finish the tile.  The builders run at most two
                            // tiles ahead of the MMAs, the MMAs at most four chunks -- four tiles when a tile is
                            // one chunk -- ahead of the reducers: 8 slots never collide)
                            sm.q2[tile_seq & 7][et] = q2;
                        }
                        fence_proxy_async_smem();
                        mbar_arrive(smem_u32(&sm.a_full[buf]));
                    }
                    chunk_seq += (uint32_t)(n_tiles * n_chunks);
                } else {
                    // ================================ reducers ====================================
                    // Two warps per TMEM lane quadrant; warp `part` of a quadrant takes the chunks whose running
                    // number has parity `part`, all 128 columns of them: load, hand the buffer back the moment
                    // the load has landed, fold.  The two warps of a scheduler wait on different barriers, so
                    // one folds while the other's load is in flight.
                    const int quad = warp & 3, part = (warp - 4) >> 2;
                    const int et = quad * 32 + lane;  // tile row = TMEM lane
                    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
                    float *dm = sm.dmin[part];
                    float m[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) m[k] = INFINITY;
                    int r0 = 0;
                    auto finish_tile = [&]() {
                        float d = min8(m);
                        if (r0 + et < rows) {
                            if (!first) d = fminf(d, dm[r0 + et]);  // minimum over the earlier model blocks
                            dm[r0 + et] = last ? d + sm.q2[tile_seq & 7][et] : d;
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) m[k] = INFINITY;
                        r0 += kSRows;
                        ++tile_seq;
                    };
                    // one chunk: wait for its MMAs, load the two 64-column halves (the second load flies under the
                    // fold of the first), hand the buffer back as soon as both have landed, fold
                    auto take = [&](uint32_t full_bar, uint32_t parity, uint32_t empty_bar, uint32_t col) {
                        uint32_t lo[64], hi[64];
                        mbar_wait(full_bar, parity);
                        tc_fence_after();
                        tmem_ld64_issue(lane_base + col, lo);
                        tmem_ld64_wait(lo);
                        tmem_ld64_issue(lane_base + col + 64u, hi);
                        fold64(lo, m);
                        tmem_ld64_wait(hi);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(empty_bar);
                        fold64(hi, m);
                    };
                    if constexpr (FIXED) {
                        // chunk c = 2 j + part of a tile: buffer c & 3, parity c >> 2 -- constants once `part` is one
                        auto tiles = [&](auto part_tag) {
                            constexpr int PART = decltype(part_tag)::value;
                            for (int tl = 0; tl < n_tiles; ++tl) {
#pragma unroll
                                for (int j = 0; j < kSFixedChunks / 2; ++j) {
                                    const uint32_t slot = (uint32_t)((2 * j + PART) & 3);
                                    take(full0 + 8u * slot, (uint32_t)((j >> 1) & 1), empty0 + 8u * slot, slot * kSChunk);
                                }
                                finish_tile();
                            }
                        };
                        if (part == 0) tiles(std::integral_constant<int, 0>{});
                        else tiles(std::integral_constant<int, 1>{});
                    } else {
                        for (int tl = 0; tl < n_tiles; ++tl) {
                            for (int c = 0; c < n_chunks; ++c, ++chunk_seq) {
                                if ((int)(chunk_seq & 1) != part) continue;
                                const uint32_t slot = chunk_seq & (kSBufs - 1);
                                take(full0 + 8u * slot, (chunk_seq / kSBufs) & 1, empty0 + 8u * slot, slot * kSChunk);
                            }
                            finish_tile();
                        }
                    }
                }
            }
            // deterministic per-hypothesis sums: one reducer warp per hypothesis, fixed order
            if (is_reducer) {
                reducer_barrier();
                for (int hh = warp - 4; hh < nb; hh += kSReducers / 32) {
                    float s = 0.f;
                    for (int i = lane; i < nfg; i += 32) {
                        const int r = hh * nfg + i;
                        s += sqrtf(fmaxf(fminf(sm.dmin[0][r], sm.dmin[1][r]), 0.f));  // clamp, then sqrt (MU:267)
                    }
                    s = warp_sum(s);
                    if (lane == 0) scores[(size_t)b * H + hb + hh] = __fdiv_rn((float)nfg, s + 1e-8f);  // MU:271
                }
                reducer_barrier();  // dmin is rewritten by the next batch
            }
        }
    } else if (is_reducer) {
        // no foreground point: sum = 0 -> score = 0 / 1e-8 = 0 (MU:271)
        for (int h = h_begin + (tid - 32 * kSIssuers - kSBuilders); h < h_end; h += kSReducers) scores[(size_t)b * H + h] = 0.f;
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

// returns PEM_E_UNSUPPORTED for shapes the tensor-core kernel does not take (the caller falls back
// to the SIMT kernel)
int score_tc_launch(const float *R, const float *t, const int32_t *sel, int B, int P, int H,
                    const float *pts1, const float *w1, int N1, const float *model, int Nm,
                    float *scores, cudaStream_t st) {
    using namespace tc;
    if (Nm > 4 * kSMaxNm || N1 > kSMaxN1) return PEM_E_UNSUPPORTED;
    const int sms = pem_num_sms();
    const int cpi = max(1, min(sms / B, (H + 7) / 8));  // CTAs per instance
    const size_t smem = sizeof(ScoreSmem) + 1024;
    const int tail = Nm % kSMaxNm;  // every model block eight chunks?
    const bool fixed = tail == 0 || tail > kSMaxNm - kSChunk;
    if (fixed) {
        static PemSmemCache cached(0);
        if (int rc = pem_ensure_dyn_smem(score_tc_kernel<true>, smem, &cached)) return rc;
        score_tc_kernel<true><<<B * cpi, kSThreads, smem, st>>>(R, t, sel, P, H, pts1, w1, N1, model, Nm, cpi, scores);
    } else {
        static PemSmemCache cached(0);
        if (int rc = pem_ensure_dyn_smem(score_tc_kernel<false>, smem, &cached)) return rc;
        score_tc_kernel<false><<<B * cpi, kSThreads, smem, st>>>(R, t, sel, P, H, pts1, w1, N1, model, Nm, cpi, scores);
    }
    return pem_launch_status();
}

}  // namespace pem
