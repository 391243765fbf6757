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
// One persistent CTA per (instance, slice of its hypotheses):
//   * the model operand (1024 rows x 64 used bytes, SWIZZLE_128B K-major rows) is built once in
//     shared memory; rows of all hypotheses of the slice are packed back to back into 128-row
//     tiles, so no tensor work is wasted on padding 151 points to 256;
//   * warps 2-9 ("row threads", two per TMEM lane = tile row, one per half of the columns)
//     transform their point, write their half of the A row (generic stores + proxy fence), and
//     later reduce their lane of the accumulator with 3-input minima straight out of TMEM
//     (PARTS = 4 puts four threads on a row; measured 6 % slower, kept as a development knob);
//   * warp 1 issues M128 x N256 x K8 MMAs: 2 per 256-model-point chunk, accumulator chunks double
//     buffered in the 512 TMEM columns; A tiles double buffered so tile t+1 is built while tile t
//     is reduced;
//   * distances land in shared memory, and each hypothesis is summed in a fixed order.
// Bound: every pair value crosses the TMEM read port once.  Measured here: 128 KB per chunk per SM in
// ~980 cycles = 128 B/clk/SM (32 B/clk per scheduler quadrant), whatever the split of the columns over
// warps (2 or 4 per row) and whether or not a warp keeps a second load in flight -- the port, not issue
// or latency, sets the pace (ncu: issue slots 37 % busy, ALU pipe 43 %, tensor pipe 30 %).
// More than 1024 model points are taken in blocks of 1024: the tiles of a batch of hypotheses run
// once per block and the running minima wait in shared memory.  The SIMT kernel (coarse.cu) remains
// for shapes this one does not take (more than 4096 model points or 256 query points).
#include "tc_ptx.cuh"

namespace pem {
namespace tc {

constexpr int kSRows = 128;          // rows per tile = TMEM lanes
constexpr int kSChunk = 256;         // model points per accumulator buffer
constexpr int kSMaxNm = 1024;        // model points held in shared memory (128 B per row)
constexpr int kSMaxN1 = 256;         // query points per instance
constexpr int kSBatchHyps = 40;      // hypotheses whose distances fit in shared memory at once
constexpr uint32_t kSIdesc = make_idesc_tf32(kSRows, kSChunk);

// PARTS = row threads per tile row, each reducing 256 / PARTS columns of every chunk (2 or 4)
template <int PARTS>
struct ScoreSmem {
    uint8_t b[kSMaxNm * 128];        // model operand
    uint8_t a[2][kSRows * 128];      // query tiles
    float dis[kSBatchHyps * kSMaxN1];
    float comb[2][PARTS - 1][kSRows];  // best of column parts 1.., handed to the part-0 thread of the row
    float px[kSMaxN1], py[kSMaxN1], pz[kSMaxN1];  // compacted foreground points
    float Rt[kSBatchHyps][12];       // R (row-major) | t of the hypotheses of the batch
    int nfg;
    uint64_t a_full[2], a_empty[2], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

template <int N>
__device__ __forceinline__ void row_barrier() {
    asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory");
}

template <int PARTS>
__global__ void __launch_bounds__(64 + 128 * PARTS, 1)
score_tc_kernel(const float *__restrict__ R, const float *__restrict__ t, const int32_t *__restrict__ sel,
                int P, int H, const float *__restrict__ pts1, const float *__restrict__ w1, int N1,
                const float *__restrict__ model, int Nm, int cpi, float *__restrict__ scores) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int kSRowThreads = 128 * PARTS, kSThreads = 64 + kSRowThreads, kSCols = kSChunk / PARTS;
    ScoreSmem<PARTS> &sm = *reinterpret_cast<ScoreSmem<PARTS> *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = warp_id(), lane = lane_id(), tid = threadIdx.x;
    const int b = blockIdx.x / cpi, slice = blockIdx.x - b * cpi;
    const int hpc = (H + cpi - 1) / cpi;                       // hypotheses per CTA
    const int h_begin = slice * hpc, h_end = min(H, h_begin + hpc);

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&sm.a_full[i]), kSRowThreads);
            mbar_init(smem_u32(&sm.a_empty[i]), 1);
            mbar_init(smem_u32(&sm.tmem_full[i]), 1);
            mbar_init(smem_u32(&sm.tmem_empty[i]), kSRowThreads / 32);
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
    // ---- per-CTA setup: foreground points (ordered compaction)
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
    const int n_blocks = (Nm + kSMaxNm - 1) / kSMaxNm;  // model blocks of <= 1024 points

    // the model operand: block `mblk` of the model points as B rows
    auto build_model = [&](int mblk) {
        const float *mb = model + (size_t)b * Nm * 3;
        const int m_lo = mblk * kSMaxNm, cnt = min(kSMaxNm, Nm - m_lo);
        const int padded = (cnt + kSChunk - 1) / kSChunk * kSChunk;
        for (int m = tid; m < padded; m += kSThreads) {
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
        return padded / kSChunk;
    };

    if (nfg > 0 && h_begin < h_end) {
        // Every warp walks the same (batch of hypotheses, model block) sequence; with more than 1024
        // model points the tiles of a batch are run once per model block and the running minimum of
        // each row waits in `dis` in between.  Barrier phases (tile_seq, chunk_seq) run on throughout.
        uint32_t tile_seq = 0, chunk_seq = 0;
        int loaded = -1;  // model block currently in shared memory
        for (int hb = h_begin; hb < h_end; hb += kSBatchHyps) {
            const int nb = min(h_end, hb + kSBatchHyps) - hb, rows = nb * nfg;
            const int n_tiles = (rows + kSRows - 1) / kSRows;
            if (warp >= 2)
                for (int k = tid - 64; k < nb * 12; k += kSRowThreads) {  // stage R | t of the batch
                    const int hh = k / 12, e = k - hh * 12, h = hb + hh;
                    const size_t src = (size_t)b * P + (sel ? __ldg(sel + (size_t)b * H + h) : h);
                    sm.Rt[hh][e] = e < 9 ? __ldg(R + src * 9 + e) : __ldg(t + src * 3 + (e - 9));
                }
            for (int mblk = 0; mblk < n_blocks; ++mblk) {
                int n_chunks = (min(kSMaxNm, Nm - mblk * kSMaxNm) + kSChunk - 1) / kSChunk;
                if (loaded != mblk) {  // (every MMA that read the previous block has retired: its
                    __syncthreads();   //  accumulators were all consumed before this barrier)
                    n_chunks = build_model(mblk);
                    loaded = mblk;
                }
                __syncthreads();
                const bool first = mblk == 0, last = mblk == n_blocks - 1;

                if (warp == 1) {
                    // ================================ MMA issuer ==================================
                    const uint64_t b0 = umma_desc_sw128(smem_u32(sm.b));
                    for (int tl = 0; tl < n_tiles; ++tl, ++tile_seq) {
                        const int buf = tile_seq & 1;
                        mbar_wait(smem_u32(&sm.a_full[buf]), (tile_seq >> 1) & 1);
                        tc_fence_after();
                        const uint64_t ad = umma_desc_sw128(smem_u32(sm.a[buf]));
                        for (int c = 0; c < n_chunks; ++c, ++chunk_seq) {
                            const int cb = chunk_seq & 1;
                            mbar_wait(smem_u32(&sm.tmem_empty[cb]), ((chunk_seq >> 1) & 1) ^ 1);
                            tc_fence_after();
                            if (elect_one()) {
                                const uint32_t d = tmem_base + (uint32_t)(cb * kSChunk);
                                const uint64_t bd = b0 + (uint64_t)(c * (kSChunk * 128 / 16));
                                tc_mma_tf32(d, ad, bd, kSIdesc, 0);          // K slots 0..7
                                tc_mma_tf32(d, ad + 2, bd + 2, kSIdesc, 1);  // K slots 8..15 (+32 bytes)
                                tc_commit(smem_u32(&sm.tmem_full[cb]));
                                if (c == n_chunks - 1) tc_commit(smem_u32(&sm.a_empty[buf]));
                            }
                            __syncwarp();
                        }
                    }
                } else if (warp >= 2) {
                    // ================================ row threads =================================
                    const int quad = warp & 3, part = (warp - 2) >> 2;
                    const int et = quad * 32 + lane;  // tile row = TMEM lane

                    // transform the point of tile row `et` and write this thread's share of the A row
                    auto build = [&](int r0, uint32_t seq) {
                        const int buf = seq & 1;
                        mbar_wait(smem_u32(&sm.a_empty[buf]), ((seq >> 1) & 1) ^ 1);  // MMAs of the tile two back retired
                        const int row = r0 + et;
                        float4 c[4];
                        c[0] = c[1] = c[2] = c[3] = make_float4(0.f, 0.f, 0.f, 0.f);
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
                            c[0] = make_float4(xh, yh, zh, xh);
                            c[1] = make_float4(yh, zh, xl, yl);
                            c[2] = make_float4(zl, xl, yl, zl);
                            c[3] = make_float4(1.f, 1.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k)  // 16 K slots = 4 float4, shared out between the parts
                            if (k * PARTS / 4 == part)
                                *reinterpret_cast<float4 *>(sm.a[buf] + sw128_offset(et, k)) = c[k];
                        fence_proxy_async_smem();
                        mbar_arrive(smem_u32(&sm.a_full[buf]));
                        return q2;
                    };

                    float q2_next = build(0, tile_seq);
                    for (int tl = 0; tl < n_tiles; ++tl, ++tile_seq) {
                        const int r0 = tl * kSRows;
                        const float q2 = q2_next;
                        if (tl + 1 < n_tiles) q2_next = build(r0 + kSRows, tile_seq + 1);
                        // reduce this thread's lane over its half of every chunk
                        float m0 = INFINITY, m1 = INFINITY, m2 = INFINITY, m3 = INFINITY;
                        for (int c = 0; c < n_chunks; ++c, ++chunk_seq) {
                            const int cb = chunk_seq & 1;
                            mbar_wait(smem_u32(&sm.tmem_full[cb]), (chunk_seq >> 1) & 1);
                            tc_fence_after();
                            {
                                float v[kSCols];
                                const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) +
                                                    (uint32_t)(cb * kSChunk + part * kSCols);
                                if constexpr (kSCols == 128) {
                                    tmem_ld128(ta, v);
                                } else {
                                    uint32_t r[64];
                                    tmem_ld64_issue(ta, r);
                                    tmem_ld64_wait(r);
#pragma unroll
                                    for (int q = 0; q < 64; ++q) v[q] = __uint_as_float(r[q]);
                                }
#pragma unroll
                                for (int q = 0; q < kSCols; q += 8) {
                                    m0 = fminf(fminf(m0, v[q]), v[q + 1]);
                                    m1 = fminf(fminf(m1, v[q + 2]), v[q + 3]);
                                    m2 = fminf(fminf(m2, v[q + 4]), v[q + 5]);
                                    m3 = fminf(fminf(m3, v[q + 6]), v[q + 7]);
                                }
                            }
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(smem_u32(&sm.tmem_empty[cb]));
                        }
                        const float best = fminf(fminf(m0, m1), fminf(m2, m3));
                        const int par = tile_seq & 1;
                        if (part > 0) sm.comb[par][part - 1][et] = best;
                        row_barrier<kSRowThreads>();
                        if (part == 0 && r0 + et < rows) {
                            float d = best;
#pragma unroll
                            for (int k = 0; k < PARTS - 1; ++k) d = fminf(d, sm.comb[par][k][et]);
                            if (!first) d = fminf(d, sm.dis[r0 + et]);  // minimum over the earlier model blocks
                            sm.dis[r0 + et] = last ? sqrtf(fmaxf(q2 + d, 0.f)) : d;  // clamp, then sqrt
                        }
                    }
                } else {  // warp 0 keeps the sequence numbers in step
                    tile_seq += (uint32_t)n_tiles;
                    chunk_seq += (uint32_t)(n_tiles * n_chunks);
                }
            }
            // deterministic per-hypothesis sums: one warp per hypothesis, fixed order
            __syncthreads();
            if (warp >= 2)
                for (int hh = warp - 2; hh < nb; hh += kSRowThreads / 32) {
                    float s = 0.f;
                    for (int i = lane; i < nfg; i += 32) s += sm.dis[hh * nfg + i];
                    s = warp_sum(s);
                    if (lane == 0) scores[(size_t)b * H + hb + hh] = __fdiv_rn((float)nfg, s + 1e-8f);  // MU:271
                }
            __syncthreads();  // dis and Rt are rewritten by the next batch
        }
    } else if (warp >= 2) {
        // no foreground point: sum = 0 -> score = 0 / 1e-8 = 0 (MU:271)
        for (int h = h_begin + (tid - 64); h < h_end; h += kSRowThreads) scores[(size_t)b * H + h] = 0.f;
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

static int g_score_parts = 2;  // row threads per tile row (development knob: 2 or 4; 4 measured 6 % slower)

// returns PEM_E_UNSUPPORTED for shapes the tensor-core kernel does not take (the caller falls back
// to the SIMT kernel)
int score_tc_launch(const float *R, const float *t, const int32_t *sel, int B, int P, int H,
                    const float *pts1, const float *w1, int N1, const float *model, int Nm,
                    float *scores, cudaStream_t st) {
    using namespace tc;
    if (Nm > 4 * kSMaxNm || N1 > kSMaxN1) return PEM_E_UNSUPPORTED;
    const int sms = pem_num_sms();
    const int cpi = max(1, min(sms / B, (H + 7) / 8));  // CTAs per instance
    if (g_score_parts == 2) {
        const size_t smem = sizeof(ScoreSmem<2>) + 1024;
        static PemSmemCache cached(0);
        if (int rc = pem_ensure_dyn_smem(score_tc_kernel<2>, smem, &cached)) return rc;
        score_tc_kernel<2><<<B * cpi, 64 + 256, smem, st>>>(R, t, sel, P, H, pts1, w1, N1, model, Nm, cpi, scores);
    } else {
        const size_t smem = sizeof(ScoreSmem<4>) + 1024;
        static PemSmemCache cached(0);
        if (int rc = pem_ensure_dyn_smem(score_tc_kernel<4>, smem, &cached)) return rc;
        score_tc_kernel<4><<<B * cpi, 64 + 512, smem, st>>>(R, t, sel, P, H, pts1, w1, N1, model, Nm, cpi, scores);
    }
    return pem_launch_status();
}

}  // namespace pem

extern "C" void pem_dev_set_score_parts(int parts) { pem::g_score_parts = parts == 4 ? 4 : 2; }
