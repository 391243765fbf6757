// Fused fine-stage path on the 5th-generation tensor cores (tcgen05 + TMEM + TMA):
// compute_feature_similarity (MU:131-153) + compute_fine_Rt (MU:308-341) WITHOUT ever
// materialising the (B, 2049, 2049) attention matrix in HBM.  (MU = the reference's
// utils/model_utils.py.)
//
//   prep      normalise rows (F.normalize, eps 1e-12) and split every fp32 value into
//             bf16 hi + bf16 lo; operands are stored row-padded to a multiple of 256.  The
//             background token (row 0 of both feature sets) stays OUT of the GEMM: its row and
//             column of the similarity matrix are 2 x 2049 exact fp32 dot products computed
//             here, so the GEMM is exactly 2048 x 2048 per instance (no 2049 -> 2304 padding).
//   GEMM      two arithmetics (PEM_PREC_*, include/pem_b200.h):
//               split-bf16  S = A_hi B_hi^T + A_hi B_lo^T + A_lo B_hi^T  (~2^-17 relative, i.e. fp32-class
//                           logits; plain bf16 would move the pose by up to 0.05 deg);
//               fp16        S = A B^T on fp16(x) -- unit vectors: 2^-12 relative per element, logits to
//                           ~1e-3 of +-1/temp, pose within 1e-4 deg of the oracle on the parity cases; a
//                           third of the MMAs, no lo operand (its shared-memory buffers extend the ring);
//             with the accumulator in TMEM, recomputed in passes whose epilogues reduce straight out of TMEM
//             (arithmetic in packed fp32 pairs, FFMA2 / FADD2):
//               pass 1  e = exp(s/temp - 1/temp);  row sums, column sums      (softmax stats)
//               pass 2  foreground labels (MU:321-322), compared in the log2 domain:
//                       row max of (2 s/temp - log colsum_j) vs the background column,
//                       column max of (2 s/temp - log rowsum_i) vs the background row (merged pass: as
//                       "does any row beat the column's background entry", a sign-bit mask per column);
//                       AND, per row, sum_j P_ij and sum_j P_ij pts2_j over EVERY column j
//               pass 3  (ragged: one 256-column strip per instance) the same sums over the columns
//                       that pass 2 found to be background, compacted into a spare operand, to be
//                       subtracted: MU:324-330 sums over foreground columns only, ~96 % of them
//             (The original form -- labels in pass 2, sums over the foreground columns in a full
//             pass 3 -- remains selectable.)  Because features are unit vectors, |s/temp| <= 1/temp,
//             so a fixed shift replaces the running max of a streaming softmax.
//   then the weighted Procrustes + pose score kernels of fine.cu.
//
// Kernel anatomy (persistent, one CTA per SM, 320 threads [fp16 mode: 576], CTAs grouped in PAIRS = clusters of 2
// driving tcgen05 cta_group::2).  A task = (instance, 256-COLUMN strip, a few 256-row steps) and
// belongs to a pair, which issues one M=256 x N=256 x K=16 MMA per K step:
//   - each CTA keeps ITS 128 columns of the strip of B' (hi+lo, 128 KB) resident in shared memory
//     (the MMA reads the peer's half over DSMEM) and streams ITS 128 rows of every step of A'
//     through a 5-stage ring of 16 KB SWIZZLE_128B chunks; its TMEM holds the 128 x 256
//     accumulator of its rows, double buffered (all 512 columns).
//   - why pairs: with independent CTAs (N = 128 resident columns) every 16 KB streamed chunk
//     feeds only 256..512 tensor cycles, i.e. 43 B/clk per SM or ~6300 B/clk chip-wide -- exactly
//     the L2 -> SM delivery cap.  A pair does twice the MMA work per streamed byte.
//   - the resident strip is replaced chunk by chunk (one full/empty barrier pair per 64-wide K
//     chunk): the next task's chunk kc is requested as soon as the last step of the current
//     task is done with chunk kc, so strip loads hide behind the remaining MMAs and tasks can
//     be short (steps per task is chosen on the host to balance the pairs).
//   (fp16 mode at C = 256 runs the FAST form of the two single-issuer roles below: eight-slot ring = two row steps,
//   every slot, barrier, parity and descriptor a constant plus one per-step offset, one elected lane walking a whole
//   step; the generic loops serve the split-bf16 mode and other C.  The issuing warp's instruction chain, not the
//   tensor pipe, bounded the generic form: tools/pass_trace.py, profiles/r2b_fine_tc_experiments.md.)
//   warp 0      TMA producer (both CTAs; completions are signalled on the LEADER's mbarriers).
//   warp 1      TMEM allocator; in the leader CTA (cluster rank 0) the single-thread tcgen05.mma
//               issuer.  tcgen05.commit multicasts slot-free / accumulator-ready / chunk-free
//               to both CTAs.
//   warps 2-9   epilogue: tcgen05.ld 32x32b (one TMEM lane = one output row per thread, two
//   (2-17)      warps per lane quadrant, 128 columns each; fp16 mode: four per quadrant, 64 columns each).  Reductions along the columns of
//               a tile are thread-local and leave as per-(row, strip-half) partials; reductions
//               along the rows accumulate in 128 registers per thread ACROSS the steps of the
//               task and are folded across lanes once per task (31-shuffle transpose reduce per
//               32 columns + a 4-quadrant combine in shared memory) into one partial plane per
//               (row split, CTA).  Validity masks are folded into the per-row / per-column
//               constants (+inf offsets).
#include <type_traits>

#include "tc_ptx.cuh"

namespace pem {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 64;  // per-CTA operand tile; BK bf16 = 128 bytes = one SWIZZLE_128B row
constexpr int TM = 2 * BM, TN = 2 * BN;     // rows per step / columns per strip of a CTA pair
constexpr int kChunkBytes = BM * BK * 2;    // 16 KB
constexpr int kStages = 5;
constexpr int kMaxKC = 4;                   // C <= 256
constexpr int kMaxStages = kStages + kMaxKC; // single-product mode: the unused b_lo buffers extend the ring to 9 slots
constexpr uint32_t kTmemCols = 2 * TN;      // double-buffered 128 x 256 accumulator: all of TMEM

enum Mode { kStore = 0, kStats = 1, kLabel = 2, kWsum = 3, kLabelW = 4 };  // kLabelW = kLabel + kWsum in one pass

struct Params {
    int B, R, Cn;          // instances, valid rows and columns of the GEMM
    int Mp, Np;            // padded rows of A' / B' per instance (multiples of 256)
    int n_msteps, n_strips; // Mp / 256 row steps, Np / 256 column strips
    int spt, n_splits;     // row steps per task, tasks per (instance, strip)
    int spt_fixed;         // 1: the caller fixed spt (PEM_FINE_STEPS_PER_TASK); the ragged launch keeps it
    int KC;                // C / 64
    int ring;              // single-product mode: ring slots (2..9)
    int single;            // 1: one fp16 product per element (operands = fp16 of the normalised rows); 0: split-bf16, three products
    float temp;            // logits = cos / temp
    float c1, c2;          // e = exp2(cos * c1 - c2)
    // kStore: element (i,j) of the GEMM goes to atten[b][i + off][j + off], row length Cn + off
    float *atten;
    int off;
    // kStats
    float *rowpart;        // (B, parts*n_strips, Mp): row sums per column part of a strip (parts = 2, or 4 in kLabelW)
    float *colpart;        // (2*n_splits, B, Np): column sums (kStats) / maxima (kLabel) per plane
    // kLabel (log2 domain)
    const float *lrs;      // (B,Mp)  log2(rowsum_i); +inf for padded rows
    const float *lcs;      // (B,Np)  log2(colsum_j); +inf for padded columns
    float *rowmaxpart;     // (B, 2*n_strips, Mp): max over the columns of the half-strip
    // kLabelW: the column labels as predicates against thr_j = 2 c bg_b[j] - log2(background row sum)
    const float *bg_b;     // (B,Np) cosine of column j with the background token (the background row)
    const float *lbg;      // (B,2)  log2 of the background row / column sums
    // kWsum
    const float4 *G;       // (B,Np): w2_j/colsum_j * (x_j, y_j, z_j, 1); zero for padding
    float *rowpart4;       // (B, 2*n_strips, 4, Mp)
    // ragged passes (the background-column correction): strips >= strip_count[b] are skipped
    const int *strip_count;
    int strip_major;       // task order: strips outermost
    int early;             // 1: the operands are older than the predecessor grid (it only produced epilogue constants), so the
                           // producer and the issuer start under its tail instead of parking in pdl_enter()
    long long *trace;      // PEM_TRACE buffer (dev builds), else null
    int dev;               // timing experiments: 1 no MMAs, 2 no streamed loads, 4 no epilogue math; on top of 1|2|4 only:
                           // 16 no ring hand-shake, 32 no accumulator hand-shake; 64 alone: the generic loops instead
                           // of the unrolled producer / issuer (A/B); 128 statistics pass without MUFU (x*x for 2^x),
                           // 256 without column sums (dev builds only).  1, 2, 4, 128, 256 keep the unrolled form
};

// Device-side timeline of CTA 0 (-DPEM_DEV_KNOBS builds, tools/pass_trace.py): role r appends (tag, clock64) pairs
// to trace[r][..]; compiled out of the shipped library.
constexpr int kTraceLen = 512;
#ifdef PEM_DEV_KNOBS
#define PEM_TRACE(role, tag)                                                          \
    do {                                                                              \
        if (p.trace && blockIdx.x == 0 && lane == 0 && trace_n < kTraceLen) {         \
            p.trace[((role)*kTraceLen + trace_n) * 2] = (tag);                        \
            p.trace[((role)*kTraceLen + trace_n) * 2 + 1] = clock64();                \
            ++trace_n;                                                                \
        }                                                                             \
    } while (0)
#else
#define PEM_TRACE(role, tag) do { } while (0)
#endif

struct Smem {
    // operand buffers first: SWIZZLE_128B needs 1024-byte alignment
    uint8_t b_hi[kMaxKC][kChunkBytes];   // resident: this CTA's 128 columns of the strip of B'
    uint8_t b_lo[kMaxKC][kChunkBytes];   // (single-product mode: no lo operand -- these are four more ring slots)
    uint8_t a[kStages][kChunkBytes];     // streamed: this CTA's 128 rows of a step of A' (hi or lo)
    float4 G[TN];                        // per-column constants of the task (kWsum)
    alignas(16) float colc[TN];          // per-column constants of the task (kLabel: -log2 colsum)
    alignas(16) float colthr[TN];        // kLabelW: the background threshold of every column
    float colred[4][TN];                 // per-quadrant column results
    uint64_t full[kMaxStages], empty[kMaxStages], b_full[kMaxKC], b_empty[kMaxKC], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

template <int THREADS>
__device__ __forceinline__ void epi_barrier() {
    asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
}

// Epilogue warps per CTA: two per TMEM lane quadrant (128 accumulator columns each).  (Four per
// quadrant were tried for the merged label + weighted-sum pass, whose epilogue does two jobs per
// element and holds the tensor pipe at 76 %: 71 us instead of 69 us -- the extra barrier traffic
// costs more than the latency hiding gains.  The kernel is written for either.)
__host__ __device__ constexpr int epi_warps(int /*mode*/) { return 8; }
// With ONE product per element (Params::single) the MMAs of a tile take a third of the time and the
// epilogue becomes the limit (measured with the MMAs switched off: 42 us statistics, 60 us merged
// pass, against 19 us of MMAs): four warps per quadrant (64 columns each) give every scheduler four
// warps to hide the TMEM-load and MUFU latencies behind.
constexpr int kEpiWarpsSingle = 16;

struct Task {
    int b, n, m0, m1;  // instance, column strip, row steps [m0, m1)
};
__device__ __forceinline__ Task decode_task(const Params &p, int task) {
    Task t;
    const int split = task % p.n_splits, bn = task / p.n_splits;
    if (p.strip_major) {  // strips outermost: the few live tasks of a ragged pass come first
        t.n = bn / p.B;
        t.b = bn - t.n * p.B;
    } else {
        t.b = bn / p.n_strips;
        t.n = bn - t.b * p.n_strips;
    }
    t.m0 = split * p.spt;
    t.m1 = min(p.n_msteps, t.m0 + p.spt);
    return t;
}

// FAST: the single-product mode at C = 256 (four K chunks) with an EIGHT-slot ring, i.e. exactly two row steps of
// streamed chunks: the slots, barriers and TMEM buffer of a step follow from its parity, so the producer's and the
// issuer's K-chunk loops unroll with every shared-memory address, descriptor and barrier parity a constant plus one
// per-step offset.  (The generic loops below cost the issuing warp ~90 dependent instructions, 460-750 clocks, per
// K chunk -- device timeline, tools/pass_trace.py -- against 512 clocks of MMAs: the issuer, not the tensor pipe,
// the loads or the epilogue, bounded the single-product passes.)
template <int MODE, int EPW, bool FAST = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EPW, 1)
fine_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const Params p) {
    const uint32_t kIdesc = p.single ? make_idesc_f16(TM, TN) : make_idesc(TM, TN);
    const int n_hl = p.single ? 1 : 2;                    // streamed chunks per (step, K chunk): hi [, lo]
    constexpr int EPT = 32 * EPW;                         // epilogue threads
    constexpr int NPW = EPW / 4;                          // column parts of a strip (one per warp of a quadrant)
    constexpr int WC = TN / NPW;                          // accumulator columns per epilogue warp

    extern __shared__ uint8_t smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = warp_id(), lane = lane_id();
    const int crank = (int)cluster_ctarank();  // 0 = leader CTA of the pair
    [[maybe_unused]] int trace_n = 0;
    // the streamed ring: a[0..4], or b_lo[0..3] + a[0..4] (contiguous) when there is no lo operand
    const int n_stages = FAST ? 8 : p.single ? min(kMaxStages, max(2, p.ring)) : kStages;
    const uint32_t ring0 = p.single ? smem_u32(sm.b_lo[0]) + (uint32_t)(kMaxStages - n_stages) * kChunkBytes
                                    : smem_u32(sm.a[0]);
    const int n_tasks = p.B * p.n_strips * p.n_splits;
    const int task0 = (int)blockIdx.x >> 1, task_step = (int)gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxStages; ++s) {
            mbar_init(smem_u32(&sm.full[s]), 1);
            mbar_init(smem_u32(&sm.empty[s]), 1);
        }
        for (int kc = 0; kc < kMaxKC; ++kc) {
            mbar_init(smem_u32(&sm.b_full[kc]), 1);
            mbar_init(smem_u32(&sm.b_empty[kc]), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(&sm.tmem_full[a]), 1);
            mbar_init(smem_u32(&sm.tmem_empty[a]), 2 * EPW);  // one arrival per epilogue warp of the pair
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM allocation is warp-collective (the same warp of both CTAs)
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&sm.tmem_base)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;
    cluster_sync_all();  // the peer's barriers and TMEM exist before anything targets them
#ifdef PEM_DEV_KNOBS
    long long cta_t0 = 0, cta_g0 = 0, cta_t1 = 0, cta_g1 = 0;
    if (p.trace && threadIdx.x == 0) { cta_t0 = clock64(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(cta_g0)); }
#endif
    // everything above is local; global memory is touched only below.  Params::early: the loads and MMAs of the
    // first two row steps run while the predecessor (a finish kernel) is still writing the per-row / per-column
    // constants that only the epilogue warps read
    if (p.early && warp < 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    else pdl_enter();
#ifdef PEM_DEV_KNOBS
    if (p.trace && threadIdx.x == 0) { cta_t1 = clock64(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(cta_g1)); }
#endif

    // Both single-issuer roles run with warp-uniform control flow (all 32 lanes loop and wait);
    // only the instruction issue is predicated on an elected lane, so addresses and
    // descriptors stay in uniform registers (no per-instruction uniformisation loops).
    if (FAST && warp == 0) {
        // ================================ TMA producer, unrolled ======================
        uint32_t it = 0, b_phase = 0;  // row steps issued so far (slot set = parity, ring phase = bit 1)
        const uint32_t full0 = smem_u32(&sm.full[0]) & kPeerBitMask, empty0 = smem_u32(&sm.empty[0]);
        const uint32_t bfull0 = smem_u32(&sm.b_full[0]) & kPeerBitMask, bempty0 = smem_u32(&sm.b_empty[0]);
        const uint32_t bdst0 = smem_u32(sm.b_hi[0]);
        for (int task = task0; task < n_tasks; task += task_step) {
            const Task t = decode_task(p, task);
            if (p.strip_count && t.n >= __ldg(p.strip_count + t.b)) continue;  // ragged pass: nothing there
            const int rowB = t.b * p.Np + t.n * TN + crank * BN;
            for (int m = t.m0; m < t.m1; ++m, ++it) {
                const int rowA = t.b * p.Mp + m * TM + crank * BM;
                const uint32_t half = it & 1u, ph = (it >> 1) & 1u;
                const uint32_t dst = ring0 + half * (4u * kChunkBytes);
                if (elect_one()) {  // one lane walks the whole step: waits, expect_tx and the bulk loads
#pragma unroll
                    for (int kc = 0; kc < 4; ++kc) {
                        if (m == t.m0) {  // replace chunk kc of the resident strip
                            mbar_wait(bempty0 + 8u * kc, b_phase ^ 1);
                            if (crank == 0) mbar_expect_tx(bfull0 + 8u * kc, 2u * kChunkBytes);
                            tma_load_2d_pair(&tmB, bfull0 + 8u * kc, bdst0 + (uint32_t)kc * kChunkBytes, kc * BK, rowB);
                        }
                        mbar_wait(empty0 + half * 32u + 8u * kc, ph ^ 1);
                        const uint32_t bar = full0 + half * 32u + 8u * kc;
#ifdef PEM_DEV_KNOBS
                        if (p.dev & 2) { if (crank == 0) mbar_arrive(bar); continue; }  // probe: no streamed loads
#endif
                        if (crank == 0) mbar_expect_tx(bar, 2u * kChunkBytes);
                        tma_load_2d_pair(&tmA, bar, dst + (uint32_t)kc * kChunkBytes, kc * BK, rowA);
                    }
                }
                __syncwarp();
            }
            b_phase ^= 1;
        }
    } else if (FAST && warp == 1) {
        // ================================ MMA issuer (leader CTA), unrolled ===========
        if (crank == 0) {
            uint32_t it = 0, b_phase = 0;
            const uint32_t a_lo0 = (ring0 >> 4) | (1u << 16), b_lo0 = (smem_u32(sm.b_hi[0]) >> 4) | (1u << 16);
            const uint32_t full0 = smem_u32(&sm.full[0]), empty0 = smem_u32(&sm.empty[0]);
            const uint32_t bfull0 = smem_u32(&sm.b_full[0]), bempty0 = smem_u32(&sm.b_empty[0]);
            const uint32_t tfull0 = smem_u32(&sm.tmem_full[0]), tempty0 = smem_u32(&sm.tmem_empty[0]);
            constexpr uint32_t idesc = make_idesc_f16(TM, TN);
            for (int task = task0; task < n_tasks; task += task_step) {
                const Task t = decode_task(p, task);
                if (p.strip_count && t.n >= __ldg(p.strip_count + t.b)) continue;
                for (int m = t.m0; m < t.m1; ++m, ++it) {
                    const uint32_t half = it & 1u, ph = (it >> 1) & 1u;
                    const bool first = m == t.m0, last = m == t.m1 - 1;
                    mbar_wait(tempty0 + 8u * half, ph ^ 1);
                    tc_fence_after();
                    const uint32_t d = tmem_base + half * (uint32_t)TN;
                    const uint32_t a_lo = a_lo0 + half * (4u * kChunkBytes >> 4);
                    if (elect_one()) {  // one lane walks the whole step: waits, MMAs and commits
#pragma unroll
                        for (int kc = 0; kc < 4; ++kc) {
                            if (first) mbar_wait(bfull0 + 8u * kc, b_phase);
                            mbar_wait(full0 + half * 32u + 8u * kc, ph);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) {  // +32 bytes per K = 16 step: +2 in the address field
                                const uint32_t ao = a_lo + (uint32_t)(kc * (kChunkBytes >> 4) + 2 * k);
                                const uint32_t bo = b_lo0 + (uint32_t)(kc * (kChunkBytes >> 4) + 2 * k);
#ifdef PEM_DEV_KNOBS
                                if (p.dev & 1) continue;  // probe: no MMAs
#endif
                                if (kc == 0 && k == 0) tc_mma_f16_pair_lo<false>(d, ao, bo, idesc);
                                else tc_mma_f16_pair_lo<true>(d, ao, bo, idesc);
                            }
                            tc_commit_pair(empty0 + half * 32u + 8u * kc);  // frees the slot once these retire
                            if (last) tc_commit_pair(bempty0 + 8u * kc);     // chunk kc of the strip may be replaced
                            if (kc == 3) tc_commit_pair(tfull0 + 8u * half); // accumulator ready
                        }
                    }
                    __syncwarp();
                }
                b_phase ^= 1;
            }
        }
    } else if (warp == 0) {
        // ================================ TMA producer ================================
        // Every completion is counted on the LEADER's full barriers (it alone arms them, for
        // both CTAs' bytes); a peer completion that lands before the leader's expect_tx only
        // drives the transaction count negative for a moment.
        int stage = 0;
        uint32_t phase = 0, b_phase = 0;
        for (int task = task0; task < n_tasks; task += task_step) {
            const Task t = decode_task(p, task);
            if (p.strip_count && t.n >= __ldg(p.strip_count + t.b)) continue;  // ragged pass: nothing there
            const int rowB = t.b * p.Np + t.n * TN + crank * BN;
            PEM_TRACE(0, 1);
            for (int m = t.m0; m < t.m1; ++m) {
                const int rowA = t.b * p.Mp + m * TM + crank * BM;
                PEM_TRACE(0, 2);
                for (int kc = 0; kc < p.KC; ++kc) {
                    if (m == t.m0) {  // replace chunk kc of the resident strip
                        mbar_wait(smem_u32(&sm.b_empty[kc]), b_phase ^ 1);
                        if (elect_one()) {
                            const uint32_t bar = smem_u32(&sm.b_full[kc]);
                            if (crank == 0) mbar_expect_tx(bar, 2u * n_hl * kChunkBytes);
                            tma_load_2d_pair(&tmB, bar & kPeerBitMask, smem_u32(sm.b_hi[kc]), kc * BK, rowB);
                            if (n_hl == 2)
                                tma_load_2d_pair(&tmB, bar & kPeerBitMask, smem_u32(sm.b_lo[kc]), kc * BK,
                                                 p.B * p.Np + rowB);
                        }
                        __syncwarp();
                    }
                    for (int hl = 0; hl < n_hl; ++hl) {
                        if (!(p.dev & 16)) mbar_wait(smem_u32(&sm.empty[stage]), phase ^ 1);
                        if (elect_one()) {
                            const uint32_t bar = smem_u32(&sm.full[stage]);
                            if (p.dev & 2) {  // timing experiment: no streamed loads
                                if (crank == 0 && !(p.dev & 16)) mbar_arrive(bar);
                            } else {
                                if (crank == 0) mbar_expect_tx(bar, 2u * kChunkBytes);
                                tma_load_2d_pair(&tmA, bar & kPeerBitMask, ring0 + (uint32_t)stage * kChunkBytes, kc * BK,
                                                 hl * p.B * p.Mp + rowA);
                            }
                        }
                        __syncwarp();
                        if (++stage == n_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
            b_phase ^= 1;
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (leader CTA) =====================
        if (crank == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, b_phase = 0, acc_phase = 0;
            // descriptor high word is constant; the low word is (address >> 4) | LBO
            const uint64_t desc_hi = (64ull << 32) | (1ull << 46) | (2ull << 61);
            auto desc = [&](uint32_t addr) { return desc_hi | (uint64_t)(((addr >> 4) & 0x3FFFu) | (1u << 16)); };
            for (int task = task0; task < n_tasks; task += task_step) {
                const Task t = decode_task(p, task);
                if (p.strip_count && t.n >= __ldg(p.strip_count + t.b)) continue;
                for (int m = t.m0; m < t.m1; ++m) {
                    PEM_TRACE(1, 10);
                    if (!(p.dev & 32)) mbar_wait(smem_u32(&sm.tmem_empty[acc]), acc_phase ^ 1);
                    PEM_TRACE(1, 11);
                    tc_fence_after();
                    const uint32_t d = tmem_base + (uint32_t)(acc * TN);
                    for (int kc = 0; kc < p.KC; ++kc) {
                        if (m == t.m0) mbar_wait(smem_u32(&sm.b_full[kc]), b_phase);
                        const uint64_t bh = desc(smem_u32(sm.b_hi[kc])), bl = desc(smem_u32(sm.b_lo[kc]));
                        // A_hi chunk: A_hi*B_hi + A_hi*B_lo (single: the only product)
                        if (!(p.dev & 16)) mbar_wait(smem_u32(&sm.full[stage]), phase);
                        PEM_TRACE(1, 12 + kc);
                        tc_fence_after();
                        uint64_t ad = desc(ring0 + (uint32_t)stage * kChunkBytes);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) {  // +32 bytes per K=16 step: +2 in the address field
                                if (p.dev & 1) break;
                                tc_mma_bf16_pair(d, ad + 2 * k, bh + 2 * k, kIdesc, (kc | k) != 0);
                                if (n_hl == 2) tc_mma_bf16_pair(d, ad + 2 * k, bl + 2 * k, kIdesc, 1);
                            }
                            if (!(p.dev & 16)) tc_commit_pair(smem_u32(&sm.empty[stage]));  // frees the slot once these retire
                            if (n_hl == 1) {
                                if (m == t.m1 - 1) tc_commit_pair(smem_u32(&sm.b_empty[kc]));
                                if (kc == p.KC - 1 && !(p.dev & 32)) tc_commit_pair(smem_u32(&sm.tmem_full[acc]));
                            }
                        }
                        __syncwarp();
                        if (++stage == n_stages) { stage = 0; phase ^= 1; }
                        if (n_hl == 1) continue;
                        // A_lo chunk: A_lo*B_hi
                        mbar_wait(smem_u32(&sm.full[stage]), phase);
                        tc_fence_after();
                        ad = desc(ring0 + (uint32_t)stage * kChunkBytes);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) {
                                if (p.dev & 1) break;
                                tc_mma_bf16_pair(d, ad + 2 * k, bh + 2 * k, kIdesc, 1);
                            }
                            tc_commit_pair(smem_u32(&sm.empty[stage]));
                            // last step of the task: chunk kc of the strip may be replaced
                            if (m == t.m1 - 1) tc_commit_pair(smem_u32(&sm.b_empty[kc]));
                            if (kc == p.KC - 1) tc_commit_pair(smem_u32(&sm.tmem_full[acc]));  // accumulator ready
                        }
                        __syncwarp();
                        if (++stage == n_stages) { stage = 0; phase ^= 1; }
                    }
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
                b_phase ^= 1;
            }
        }
    } else {
        // ================================ epilogue ====================================
        const int quad = warp & 3;                 // TMEM lanes [32*quad, 32*quad+32)
        const int half = (warp - 2) >> 2;          // column part: columns [WC*half, WC*half+WC) of the tile
        const int et = quad * 32 + lane;           // row of this CTA's 128 rows owned by this thread
        const int ep_tid = (warp - 2) * 32 + lane; // 0..EPT-1
        const float kc1 = p.c1, kc2 = p.c2, kc2x = 2.0f * p.c1, ktemp = p.temp;
        constexpr bool LAB = (MODE == kLabel || MODE == kLabelW);  // label maxima
        constexpr bool WS = (MODE == kWsum || MODE == kLabelW);    // weighted row sums
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int task = task0; task < n_tasks; task += task_step) {
            const Task t = decode_task(p, task);
            if (p.strip_count && t.n >= __ldg(p.strip_count + t.b)) continue;
            const int b = t.b, n = t.n;
            [[maybe_unused]] const int trole = warp == 2 ? 2 : 3;
            if (warp == 2 || warp == 1 + EPW) PEM_TRACE(trole, 40);
            const int j0 = n * TN + half * WC;     // first column of this warp
            const int ncols = min(WC, p.Cn - j0);  // valid columns (<= 0: none)
            // Column reductions.  kStats / kLabel keep one accumulator per column in registers across the
            // steps of the task (cheapest in instructions).  kLabelW has no registers to spare for that:
            // it folds each 32-column chunk across the warp's 32 rows at once (31-shuffle transpose
            // reduce) and keeps only lane L's column of every chunk.
            // The arithmetic is issued as fp32 PAIRS (FFMA2 / FADD2, tc_ptx.cuh) over adjacent columns:
            // the epilogue is what bounds the single-product passes, and it is bound by issue slots.
            constexpr bool COLREG = MODE == kStats || MODE == kLabel;
            float colacc[MODE == kLabel ? WC : 1];          // kLabel: column maxima
            uint64_t colacc2[MODE == kStats ? WC / 2 : 1];  // kStats: column sums, two columns per register pair
            uint32_t cmask[WC / 32];  // kLabelW: bit (31 - q) = some row so far beats the background in column q of the chunk
#pragma unroll
            for (int q = 0; q < WC / 32; ++q) cmask[q] = 0u;
#pragma unroll
            for (int q = 0; q < (MODE == kLabel ? WC : 1); ++q) colacc[q] = -INFINITY;
#pragma unroll
            for (int q = 0; q < (MODE == kStats ? WC / 2 : 1); ++q) colacc2[q] = 0ull;
            // per-column constants of the task.  colc = -log2 colsum_j (-inf keeps padded columns out of
            // the row max).  G is stored by column PAIRS: G[2k] = (w, w', x, x'), G[2k+1] = (y, y', z, z')
            // of columns 2k and 2k+1, so that one LDS.128 feeds two FFMA2.
            if (LAB && ep_tid < TN) sm.colc[ep_tid] = -p.lcs[(size_t)b * p.Np + n * TN + ep_tid];
            if (MODE == kLabelW && ep_tid < TN)  // exactly tc_finish_label_kernel's threshold (label2 > 0)
                sm.colthr[ep_tid] = fmaf(p.bg_b[(size_t)b * p.Np + n * TN + ep_tid], kc2x, -p.lbg[2 * b]);
            if (WS && ep_tid < TN) {
                const float4 g = p.G[(size_t)b * p.Np + n * TN + ep_tid];
                float *gp = reinterpret_cast<float *>(&sm.G[ep_tid & ~1]) + (ep_tid & 1);
                gp[0] = g.w; gp[2] = g.x; gp[4] = g.y; gp[6] = g.z;
            }
            if (LAB || WS) epi_barrier<EPT>();

            for (int m = t.m0; m < t.m1; ++m) {
                const int gi = m * TM + crank * BM + et;
                const bool row_ok = gi < p.R;
                float rowc = 0.f;  // per-row constant of this tile
                if (MODE == kStats) rowc = row_ok ? kc2 : INFINITY;  // e = 0 for padded rows
                if (LAB) rowc = p.lrs[(size_t)b * p.Mp + gi];  // +inf for padded rows
                float rm = -INFINITY;                           // label: row maximum
                // statistics: two pairs of row sums; weighted sums: (even, odd column) pairs of w, x, y, z
                uint64_t a0 = 0ull, a1 = 0ull, a2 = 0ull, a3 = 0ull;
                const uint64_t K1 = pack2(kc1, kc1), K2X = pack2(kc2x, kc2x), NROW = pack2(-rowc, -rowc),
                               NC2 = pack2(-2.0f * kc2, -2.0f * kc2);
                if (warp == 2 || warp == 1 + EPW) PEM_TRACE(trole, 30);
                if (!(p.dev & 32)) mbar_wait(smem_u32(&sm.tmem_full[acc]), acc_phase);
                if (warp == 2 || warp == 1 + EPW) PEM_TRACE(trole, 31);
                tc_fence_after();
                // columns per TMEM load: 16 where 128 column accumulators already fill the registers, and in the FAST
                // form, which keeps the next chunk's load in flight under the arithmetic of the current one
                constexpr int CW = (COLREG || FAST) ? 16 : 32;
                const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * TN + half * WC);
                float v[CW];  // the chunk being worked on
                auto chunk_m = [&](const int c, auto mask_tag) {  // (c is a constant after unrolling)
                    [[maybe_unused]] constexpr bool MASK = decltype(mask_tag)::value;  // kStats: the strip has padded columns
                    if (MODE == kStore) {
                        if (row_ok) {
                            float *dst = p.atten + ((size_t)b * (p.R + p.off) + gi + p.off) * (p.Cn + p.off) +
                                         p.off + j0 + c * CW;
#pragma unroll
                            for (int q = 0; q < CW; ++q)
                                if (c * CW + q < ncols) dst[q] = __fdiv_rn(v[q], ktemp);
                        }
                    } else if (MODE == kStats) {
                        // two pairs at a time: the adds do not wait on each other.  MASK: the last strip's
                        // padded columns contribute nothing (a separate copy: predicated-off selects would
                        // still take issue slots on every other strip)
                        {
#pragma unroll
                            for (int q = 0; q < CW; q += 4) {
                                float e0, e1, e2, e3;
                                unpack2(fma2(pack2(v[q], v[q + 1]), K1, NROW), e0, e1);
                                unpack2(fma2(pack2(v[q + 2], v[q + 3]), K1, NROW), e2, e3);
#ifdef PEM_DEV_KNOBS
                                if (p.dev & 128) { e0 *= e0; e1 *= e1; e2 *= e2; e3 *= e3; }  // probe: no MUFU
                                else
#endif
                                { e0 = ex2(e0); e1 = ex2(e1); e2 = ex2(e2); e3 = ex2(e3); }
                                if (MASK) {
                                    if (c * CW + q >= ncols) e0 = 0.f;
                                    if (c * CW + q + 1 >= ncols) e1 = 0.f;
                                    if (c * CW + q + 2 >= ncols) e2 = 0.f;
                                    if (c * CW + q + 3 >= ncols) e3 = 0.f;
                                }
                                const uint64_t p01 = pack2(e0, e1), p23 = pack2(e2, e3);
                                a0 = add2(a0, p01);
                                a1 = add2(a1, p23);
#ifdef PEM_DEV_KNOBS
                                if (p.dev & 256) continue;  // probe: no column sums
#endif
                                colacc2[(c * CW + q) / 2] = add2(colacc2[(c * CW + q) / 2], p01);
                                colacc2[(c * CW + q) / 2 + 1] = add2(colacc2[(c * CW + q) / 2 + 1], p23);
                            }
                        }
                    }
                    if (LAB) {
#pragma unroll
                        for (int q = 0; q < CW; q += 4) {
                            const float4 cc = lds_f4(&sm.colc[half * WC + c * CW + q]);  // -log2 colsum_j
                            // 2 s/temp - log colsum_j (row side), 2 s/temp - log rowsum_i (column side)
                            float t0, t1, t2, t3;
                            unpack2(fma2(pack2(v[q], v[q + 1]), K2X, pack2(cc.x, cc.y)), t0, t1);
                            unpack2(fma2(pack2(v[q + 2], v[q + 3]), K2X, pack2(cc.z, cc.w)), t2, t3);
                            rm = fmaxf(fmaxf(rm, t0), t1);
                            rm = fmaxf(fmaxf(rm, t2), t3);
                            if (MODE == kLabel) {
#pragma unroll
                                for (int u = 0; u < 4; ++u)
                                    colacc[c * CW + q + u] = fmaxf(colacc[c * CW + q + u], fmaf(v[q + u], kc2x, -rowc));
                            }
                        }
                    }
                    if (WS) {
#pragma unroll
                        for (int q = 0; q < CW; q += 2) {
                            float u0, u1;
                            unpack2(fma2(pack2(v[q], v[q + 1]), K2X, NC2), u0, u1);  // e^2
                            const uint64_t u = pack2(ex2(u0), ex2(u1));
                            const float4 gA = lds_f4(&sm.G[half * WC + c * CW + q]);      // (w, w', x, x')
                            const float4 gB = lds_f4(&sm.G[half * WC + c * CW + q + 1]);  // (y, y', z, z')
                            a0 = fma2(u, pack2(gA.x, gA.y), a0);
                            a1 = fma2(u, pack2(gA.z, gA.w), a1);
                            a2 = fma2(u, pack2(gB.x, gB.y), a2);
                            a3 = fma2(u, pack2(gB.z, gB.w), a3);
                        }
                    }
                    if constexpr (LAB && !COLREG) {
                        // Column side of the labels.  Only "does any row beat the column's background entry"
                        // is needed, and that threshold is known before the pass: thr_j - val_ij < 0 with
                        // val = 2 s/temp - log rowsum_i, evaluated as fma(v, -k, +rowc) = -val (exact mirror of
                        // the fma the maximum form uses) and a rounded difference, whose SIGN is exact -- the
                        // same decision as max_i val_ij > thr_j, bit for bit.  The sign bits are shifted into a
                        // mask (one SHF per element) instead of a 31-shuffle maximum per chunk and step.
                        const uint64_t NK2X = pack2(-kc2x, -kc2x), PROW = pack2(rowc, rowc);
                        uint32_t m = 0u;
#pragma unroll
                        for (int q = 0; q < CW; q += 4) {
                            const float4 th = lds_f4(&sm.colthr[half * WC + c * CW + q]);
                            float d0, d1, d2, d3;
                            unpack2(add2(pack2(th.x, th.y), fma2(pack2(v[q], v[q + 1]), NK2X, PROW)), d0, d1);
                            unpack2(add2(pack2(th.z, th.w), fma2(pack2(v[q + 2], v[q + 3]), NK2X, PROW)), d2, d3);
                            m = shift_in_sign(m, d0);
                            m = shift_in_sign(m, d1);
                            m = shift_in_sign(m, d2);
                            m = shift_in_sign(m, d3);
                        }
                        if (CW == 32) cmask[c] |= m;
                        else cmask[c / 2] |= m << (16 * (1 - (c & 1)));  // 16-column chunks: column q of the chunk at bit 15 - q
                    }
                };
                auto chunk = [&](const int c) {
                    if (MODE == kStats && ncols < WC) chunk_m(c, std::true_type{});
                    else chunk_m(c, std::false_type{});
                };
                // accumulator fully read: hand the TMEM buffer back to the MMA warp of the leader
                auto hand_back = [&]() {
                    if (warp == 2 || warp == 1 + EPW) PEM_TRACE(trole, 32);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0 && !(p.dev & 32)) mbar_arrive_remote(smem_u32(&sm.tmem_empty[acc]), 0);
                };
                if constexpr (FAST && !COLREG) {
                    // Software pipeline over the four 16-column chunks of this warp: wait for chunk c, put the load of
                    // chunk c + 1 in flight, then do the arithmetic of chunk c; the buffer goes back as soon as the last
                    // chunk has landed in registers.  Worth 2 us on the label + weighted-sum pass and 4 us on a full
                    // weighted-sum pass.  Not for the passes that keep one accumulator per column (COLREG): 64 of them
                    // leave no room for a second chunk, and eight pipelined 8-column chunks measured slower (43.5 us)
                    // than four plain 16-column ones (40 us).
                    constexpr int NCH = WC / CW;
                    if (!(p.dev & 4)) {
                        uint32_t r[2][CW];
                        tmem_ld_issue(taddr0, r[0]);
#pragma unroll
                        for (int c = 0; c < NCH; ++c) {
                            tmem_ld_wait(r[c & 1]);
                            if (c + 1 < NCH) tmem_ld_issue(taddr0 + (uint32_t)((c + 1) * CW), r[(c + 1) & 1]);
                            else hand_back();
#pragma unroll
                            for (int q = 0; q < CW; ++q) v[q] = __uint_as_float(r[c & 1][q]);
                            chunk(c);
                        }
                    } else {
                        hand_back();
                    }
                } else {
                    if (!(p.dev & 4)) {
#pragma unroll
                        for (int c = 0; c < WC / CW; ++c) {
                            tmem_ld(taddr0 + (uint32_t)(c * CW), v);
                            chunk(c);
                        }
                    }
                    hand_back();
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                // per-(row, half-strip) partials: consecutive lanes = consecutive rows
                const size_t pr = ((size_t)b * NPW * p.n_strips + NPW * n + half);
                if (MODE == kStats) p.rowpart[pr * p.Mp + gi] = sum2(a0) + sum2(a1);
                if (LAB) p.rowmaxpart[pr * p.Mp + gi] = rm;
                if (WS) {
                    float *dst = p.rowpart4 + pr * 4 * p.Mp + gi;
                    dst[0] = sum2(a0); dst[p.Mp] = sum2(a1); dst[2 * p.Mp] = sum2(a2); dst[3 * p.Mp] = sum2(a3);
                }
            }
            // fold the column accumulators across the 32 lanes, then across the 4 quadrants;
            // every (row split, CTA of the pair) leaves its own plane
            if (warp == 2 || warp == 1 + EPW) PEM_TRACE(trole, 41);
            if (MODE == kStats || LAB) {
#pragma unroll
                for (int c = 0; c < WC / 32; ++c) {
                    if (COLREG) {
                        float v[32];
#pragma unroll
                        for (int q = 0; q < 32; q += 2) {
                            if (MODE == kStats) unpack2(colacc2[(c * 32 + q) / 2], v[q], v[q + 1]);
                            else { v[q] = colacc[c * 32 + q]; v[q + 1] = colacc[c * 32 + q + 1]; }
                        }
                        sm.colred[quad][half * WC + c * 32 + lane] = warp_transpose_reduce<LAB>(v);
                    } else {  // any row of the warp; column c*32 + L sits at bit 31 - L
                        const uint32_t any = __reduce_or_sync(kFull, cmask[c]);
                        sm.colred[quad][half * WC + c * 32 + lane] = (any >> (31 - lane)) & 1u ? INFINITY : -INFINITY;
                    }
                }
                epi_barrier<EPT>();
                if (ep_tid < TN) {
                    float s = sm.colred[0][ep_tid];
#pragma unroll
                    for (int w = 1; w < 4; ++w)
                        s = LAB ? fmaxf(s, sm.colred[w][ep_tid]) : s + sm.colred[w][ep_tid];
                    const int plane = 2 * (t.m0 / p.spt) + crank;
                    p.colpart[((size_t)plane * p.B + b) * p.Np + n * TN + ep_tid] = s;
                }
            }
            epi_barrier<EPT>();  // shared per-task constants / colred are reused by the next task
            if (warp == 2 || warp == 1 + EPW) PEM_TRACE(trole, 42);
        }
    }

    tc_fence_before();
    __syncthreads();
#ifdef PEM_DEV_KNOBS
    if (p.trace && threadIdx.x == 0) {  // per-CTA span: (clock, globaltimer) before / after the dependency wait and at the end
        long long *q = p.trace + 4 * kTraceLen * 2 + 6 * blockIdx.x, g2;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g2));
        q[0] = cta_t0; q[1] = cta_g0; q[2] = cta_t1; q[3] = cta_g1; q[4] = clock64(); q[5] = g2;
    }
#endif
    cluster_sync_all();  // the leader's MMAs read the peer's shared memory and write its TMEM
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols)
                     : "memory");
    }
}

// ---------------------------------------------------------------------------- prep / finish
// One warp per PAIR of padded rows of the GEMM operands: x / max(|x|,1e-12) split into bf16 hi and
// lo.  The first `skip` rows of every instance (the background token) stay out of the operands; with
// skip = 1 the warp also leaves the fp32 cosine of its rows with the OTHER set's background
// token, i.e. the background column (bg_a) / row (bg_b) of the similarity matrix, and their
// corner s00.  C <= 256 and C % 64 == 0: a lane holds up to NQ float4 of each row in registers
// (16-byte loads, 8-byte bf16x4 stores, every access of the warp one contiguous run; both rows'
// loads are in flight together).  rows_pa and rows_pb are even.
constexpr int kPrepRows = 2;
// Persistent: every warp walks the row pairs with a stride and loads the NEXT pair's rows before it
// works on the current one, so the global loads of one pair fly under the arithmetic and the stores
// of the previous one (a warp per pair that loads, reduces, stores and exits ran at 3.3 TB/s).
template <int NQ>  // ceil(C / 128): float4 per lane and row
__global__ void __launch_bounds__(256, 3)
tc_prep_kernel(const float *__restrict__ fa, int rows_a, int rows_pa, __nv_bfloat16 *__restrict__ outa,
               const float *__restrict__ fb, int rows_b, int rows_pb, __nv_bfloat16 *__restrict__ outb,
               int B, int C, int normalize, int skip, float *__restrict__ bg_a /* (B,rows_pa) */,
               float *__restrict__ bg_b /* (B,rows_pb) */, float *__restrict__ s00 /* (B) */,
               int single /* 1: the hi plane holds fp16(x), no lo plane is written */) {
    pdl_enter();
    const int lane = lane_id();
    const int nq = C >> 2;  // float4 per row
    const int n_warps = gridDim.x * (blockDim.x >> 5);
    const int total = B * (rows_pa + rows_pb) / kPrepRows;  // row pairs: A' first, then B'
    // every warp owns a CONTIGUOUS range of pairs: the instance (and with it the other set's background
    // token) changes at most once or twice per warp, and the position advances without divisions
    const int per = (total + n_warps - 1) / n_warps;
    const int gw = blockIdx.x * (blockDim.x >> 5) + warp_id();
    const int p_begin = gw * per, p_end = min(total, p_begin + per);
    if (p_begin >= p_end) return;
    // where a pair lives: operand, instance, first row
    struct Where { bool second; int b, i0; };
    auto locate = [&](int pair) {
        Where w;
        int gr = pair * kPrepRows;
        w.second = gr >= B * rows_pa;
        if (w.second) gr -= B * rows_pa;
        const int rows_p = w.second ? rows_pb : rows_pa;
        w.b = gr / rows_p;
        w.i0 = gr - w.b * rows_p;
        return w;
    };
    auto advance = [&](Where &w) {
        w.i0 += kPrepRows;
        if (w.i0 >= (w.second ? rows_pb : rows_pa)) {
            w.i0 = 0;
            if (++w.b == B && !w.second) { w.b = 0; w.second = true; }
        }
    };
    // raw loads of a pair's rows (zero beyond the row / for padding rows); nothing waits on them here
    auto fetch = [&](const Where &w, float4 (&x)[kPrepRows][NQ]) {
        const float *f = w.second ? fb : fa;
        const int rows = w.second ? rows_b : rows_a;
#pragma unroll
        for (int r = 0; r < kPrepRows; ++r) {
            const bool live = w.i0 + r < rows;
            const float *row = f + ((size_t)w.b * (rows + skip) + skip + w.i0 + r) * C;
#pragma unroll
            for (int it = 0; it < NQ; ++it) {
                const int q = it * 32 + lane;
                x[r][it] = (live && q < nq) ? __ldg(reinterpret_cast<const float4 *>(row) + q)
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };
    // arithmetic in fp32 pairs (FFMA2 / FMUL2): the kernel was issue bound (266 warp instructions per row)
    auto sumsq = [&](const float4 (&x)[NQ]) {
        uint64_t s = 0ull;
#pragma unroll
        for (int it = 0; it < NQ; ++it) {
            const uint64_t lo = pack2(x[it].x, x[it].y), hi = pack2(x[it].z, x[it].w);
            s = fma2(lo, lo, s);
            s = fma2(hi, hi, s);
        }
        return sum2(s);
    };
    auto dot_of = [&](const float4 (&x)[NQ], const float4 (&y)[NQ]) {
        uint64_t s = 0ull;
#pragma unroll
        for (int it = 0; it < NQ; ++it) {
            s = fma2(pack2(x[it].x, x[it].y), pack2(y[it].x, y[it].y), s);
            s = fma2(pack2(x[it].z, x[it].w), pack2(y[it].z, y[it].w), s);
        }
        return sum2(s);
    };
    // a row as NQ float4 per lane (zero beyond the row); sum of squares (unreduced)
    auto load_row = [&](const float *row, float4 (&x)[NQ]) {
#pragma unroll
        for (int it = 0; it < NQ; ++it) {
            const int q = it * 32 + lane;
            x[it] = q < nq ? __ldg(reinterpret_cast<const float4 *>(row) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return sumsq(x);
    };
    auto inv_of = [&](float ss) { return normalize ? fmaxf(sqrtf(warp_sum(ss)), 1e-12f) : 1.f; };
    auto split_store = [&](float4 v, uint2 *hi, uint2 *lo, int q) {  // 4 values -> bf16 hi / lo, 8 bytes each
        if (single) {  // one fp16 value per element (|x| <= 1: no range problem, 2^-12 relative)
            const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
            hi[q] = make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
            return;
        }
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
        const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
        const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y),
                             l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
        hi[q] = make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
        lo[q] = make_uint2(*reinterpret_cast<const uint32_t *>(&l0), *reinterpret_cast<const uint32_t *>(&l1));
    };

    float4 x[kPrepRows][NQ], xn[kPrepRows][NQ], tk[NQ];
    Where w = locate(p_begin), wn = w;
    fetch(w, x);
    bool tk_second = false;
    int tk_b = -1;
    float rinv_tok = 1.f;  // 1 / |token|
    for (int pair = p_begin; pair < p_end; ++pair) {
        const bool more = pair + 1 < p_end;
        if (more) {
            advance(wn);
            fetch(wn, xn);
        }
        const bool second = w.second;
        const int b = w.b, i0 = w.i0;  // rows i0, i0+1 of instance b
        const float *f = second ? fb : fa, *fo = second ? fa : fb;
        const int rows = second ? rows_b : rows_a, rows_p = second ? rows_pb : rows_pa;
        const int rows_o = second ? rows_a : rows_b;
        const int gr = b * rows_p + i0;
        __nv_bfloat16 *out = second ? outb : outa;  // (2, B*rows_p, C)
        float *bg = second ? bg_b : bg_a;
        if (skip && (tk_b != b || tk_second != second)) {  // the other set's background token of this instance
            rinv_tok = 1.0f / inv_of(load_row(fo + (size_t)b * (rows_o + skip) * C, tk));
            tk_b = b;
            tk_second = second;
        }
#pragma unroll
        for (int r = 0; r < kPrepRows; ++r) {
            const int i = i0 + r;
            uint2 *hi = reinterpret_cast<uint2 *>(out + (size_t)(gr + r) * C),
                  *lo = reinterpret_cast<uint2 *>(out + ((size_t)B * rows_p + gr + r) * C);
            // (padding rows, i >= rows, were fetched as zeros and come out as zeros below -- operand planes
            // and background cosine alike -- so the row loop stays free of divergent branches around its
            // warp reductions)
            // x * (1/norm) instead of the reference's x / norm: at most 1 ulp apart, far below the
            // 2^-12 / 2^-17 of the rounding that follows, and 8 IEEE divisions per lane cheaper
            const float rinv = normalize ? 1.0f / inv_of(sumsq(x[r])) : 1.f;
            const uint64_t rinv2 = pack2(rinv, rinv);
#pragma unroll
            for (int it = 0; it < NQ; ++it) {
                float4 &v = x[r][it];
                unpack2(mul2(pack2(v.x, v.y), rinv2), v.x, v.y);
                unpack2(mul2(pack2(v.z, v.w), rinv2), v.z, v.w);
                const int q = it * 32 + lane;
                if (q < nq) split_store(v, hi, lo, q);
            }
            if (!skip) continue;
            const float dot = warp_sum(dot_of(x[r], tk)) * rinv_tok;  // cosine with the background token
            if (lane == 0) bg[(size_t)b * rows_p + i] = dot;
            if (!second && i == 0) {  // the corner: background token x background token
                float4 own[NQ];
                const float inv_own = inv_of(load_row(f + (size_t)b * (rows + skip) * C, own));
                const float d0 = __fdiv_rn(warp_sum(dot_of(own, tk)), inv_own) * rinv_tok;
                if (lane == 0) s00[b] = d0;
            }
        }
        if (more) {
            w = wn;
#pragma unroll
            for (int r = 0; r < kPrepRows; ++r)
#pragma unroll
                for (int it = 0; it < NQ; ++it) x[r][it] = xn[r][it];
        }
    }
}

// kStore only: the background row and column of the logits (the GEMM writes the interior)
__global__ void __launch_bounds__(256)
tc_store_bg_kernel(const float *__restrict__ bg_a, const float *__restrict__ bg_b,
                   const float *__restrict__ s00, int R, int Cn, int Mp, int Np, float temp,
                   float *__restrict__ atten /* (B,R+1,Cn+1) */) {
    const int b = blockIdx.y, k = blockIdx.x * blockDim.x + threadIdx.x;
    float *dst = atten + (size_t)b * (R + 1) * (Cn + 1);
    if (k == 0) dst[0] = __fdiv_rn(s00[b], temp);
    if (k < R) dst[(size_t)(k + 1) * (Cn + 1)] = __fdiv_rn(bg_a[(size_t)b * Mp + k], temp);
    if (k < Cn) dst[k + 1] = __fdiv_rn(bg_b[(size_t)b * Np + k], temp);
}

// The finish kernels fold per-(row, column part) and per-(column, plane) partials.  A block of 256
// threads serves 64 rows / columns with FOUR sub-threads each (thread = sub * 64 + item: loads stay
// coalesced along the rows): a sub-thread folds every fourth partial with its loads in flight
// together, the four results meet in shared memory in a fixed order (deterministic).  One thread per
// row with a serial loop over 32..128 partials was latency bound (5-9 us per kernel).
constexpr int kFinishItems = 64, kFinishSub = 4;
template <bool MAX>
__device__ __forceinline__ float part_fold(const float *__restrict__ base, size_t stride, int n, int sub) {
    float a = MAX ? -INFINITY : 0.f;
#pragma unroll 8
    for (int t = sub; t < n; t += kFinishSub) {
        const float v = base[(size_t)t * stride];
        a = MAX ? fmaxf(a, v) : a + v;
    }
    return a;
}
// result valid in sub-thread 0; `red` is reused by consecutive calls (barriers inside)
template <bool MAX>
__device__ __forceinline__ float sub_fold(float v, float (*red)[kFinishItems], int item, int sub) {
    __syncthreads();
    red[sub][item] = v;
    __syncthreads();
    if (sub != 0) return v;
    float a = red[0][item];
#pragma unroll
    for (int q = 1; q < kFinishSub; ++q) a = MAX ? fmaxf(a, red[q][item]) : a + red[q][item];
    return a;
}

// after pass 1: row / column sums from their partials (fixed order: deterministic) plus the
// background column / row term, reciprocals and log2; +inf logs mark padded rows / columns
__global__ void __launch_bounds__(256)
tc_finish_stats_kernel(const float *__restrict__ rowpart, const float *__restrict__ colpart,
                       const float *__restrict__ bg_a, const float *__restrict__ bg_b, int R, int Cn,
                       const float *__restrict__ s00, int Mp, int Np, int n_part, int planes, float c1,
                       float c2, float *__restrict__ inv_rowsum, float *__restrict__ lrs,
                       float *__restrict__ inv_colsum, float *__restrict__ lcs,
                       float *__restrict__ lbg /* (B,2): log2 of the background row / column sums */,
                       const float *__restrict__ pts2 /* (B,Cn,3) or null */,
                       float4 *__restrict__ Gall /* (B,Np) or null: 1/colsum_j * (x_j, y_j, z_j, 1), EVERY column */) {
    __shared__ float red[32];
    __shared__ float red4[kFinishSub][kFinishItems];
    pdl_enter();
    const int b = blockIdx.y, B = gridDim.y;
    const int item = threadIdx.x % kFinishItems, sub = threadIdx.x / kFinishItems;
    const int k = blockIdx.x * kFinishItems + item;
    if (blockIdx.x == gridDim.x - 1) {  // one EXTRA block per instance: sums of the background row and column
        float sr = 0.f, sc = 0.f;
        for (int j = threadIdx.x; j < Cn; j += blockDim.x) sr += ex2(fmaf(bg_b[(size_t)b * Np + j], c1, -c2));
        for (int i = threadIdx.x; i < R; i += blockDim.x) sc += ex2(fmaf(bg_a[(size_t)b * Mp + i], c1, -c2));
        const float e00 = ex2(fmaf(s00[b], c1, -c2));
        sr = block_sum(sr, red);
        sc = block_sum(sc, red);
        if (threadIdx.x == 0) { lbg[2 * b] = log2f(sr + e00); lbg[2 * b + 1] = log2f(sc + e00); }
        return;
    }
    // rows: the partials of row k are shared by the four sub-threads, then folded in a fixed order
    {
        float s = 0.f;
        if (k < R) s = part_fold<false>(rowpart + (size_t)b * n_part * Mp + k, Mp, n_part, sub);
        s = sub_fold<false>(s, red4, item, sub);
        if (sub == 0 && k < Mp) {
            if (k < R) s += ex2(fmaf(bg_a[(size_t)b * Mp + k], c1, -c2));
            const bool ok = k < R && s > 0.f;
            inv_rowsum[(size_t)b * Mp + k] = ok ? 1.0f / s : 0.f;
            lrs[(size_t)b * Mp + k] = ok ? log2f(s) : INFINITY;
        }
    }
    {
        float s = 0.f;
        if (k < Cn) s = part_fold<false>(colpart + (size_t)b * Np + k, (size_t)B * Np, planes, sub);
        s = sub_fold<false>(s, red4, item, sub);
        if (sub == 0 && k < Np) {
            if (k < Cn) s += ex2(fmaf(bg_b[(size_t)b * Np + k], c1, -c2));
            const bool ok = k < Cn && s > 0.f;
            const float ic = ok ? 1.0f / s : 0.f;
            inv_colsum[(size_t)b * Np + k] = ic;
            lcs[(size_t)b * Np + k] = ok ? log2f(s) : INFINITY;
            if (Gall) {
                float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < Cn) {
                    const float *q = pts2 + ((size_t)b * Cn + k) * 3;
                    g = make_float4(ic * q[0], ic * q[1], ic * q[2], ic);
                }
                Gall[(size_t)b * Np + k] = g;
            }
        }
    }
}

// after pass 2: w1_i, w2_j and the per-column constants of pass 3.  A point is foreground when its
// best match beats the background entry (MU:321-322); the background row / column sums are reduced
// by tc_finish_stats_kernel from the 2 x 2049 exact cosines of the prep kernel.
__global__ void __launch_bounds__(256)
tc_finish_label_kernel(const float *__restrict__ rowmaxpart, const float *__restrict__ colpart,
                       const float *__restrict__ bg_a, const float *__restrict__ bg_b,
                       const float *__restrict__ lbg, const float *__restrict__ inv_colsum,
                       const float *__restrict__ pts2, int R, int Cn, int Mp, int Np, int n_part,
                       int planes, float c1, float c2, float *__restrict__ w1 /* (B,R) */,
                       float *__restrict__ w2_out /* (B,Cn) or null */, float4 *__restrict__ G /* or null */,
                       float *__restrict__ fgcol /* (B,Np) or null: 1 = foreground or padding, 0 = background */) {
    __shared__ float red4[kFinishSub][kFinishItems];
    pdl_enter();
    const int b = blockIdx.y, B = gridDim.y;
    const int item = threadIdx.x % kFinishItems, sub = threadIdx.x / kFinishItems;
    const int k = blockIdx.x * kFinishItems + item;
    const float lrs0 = lbg[2 * b], lcs0 = lbg[2 * b + 1], c2x = 2.0f * c1;
    {
        float m = -INFINITY;
        if (k < R) m = part_fold<true>(rowmaxpart + (size_t)b * n_part * Mp + k, Mp, n_part, sub);
        m = sub_fold<true>(m, red4, item, sub);
        if (sub == 0 && k < R)
            w1[(size_t)b * R + k] = m > fmaf(bg_a[(size_t)b * Mp + k], c2x, -lcs0) ? 1.f : 0.f;  // label1 > 0
    }
    {
        float cm = -INFINITY;
        if (k < Cn) cm = part_fold<true>(colpart + (size_t)b * Np + k, (size_t)B * Np, planes, sub);
        cm = sub_fold<true>(cm, red4, item, sub);
        if (sub == 0 && k < Np) {
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            bool isbg = false;
            if (k < Cn) {
                const bool fg = cm > fmaf(bg_b[(size_t)b * Np + k], c2x, -lrs0);  // label2 > 0
                isbg = !fg;
                if (w2_out) w2_out[(size_t)b * Cn + k] = fg ? 1.f : 0.f;
                if (fg) {
                    const float ic = inv_colsum[(size_t)b * Np + k];
                    const float *q = pts2 + ((size_t)b * Cn + k) * 3;
                    g = make_float4(ic * q[0], ic * q[1], ic * q[2], ic);
                }
            }
            if (G) G[(size_t)b * Np + k] = g;
            if (fgcol) fgcol[(size_t)b * Np + k] = isbg ? 0.f : 1.f;
        }
    }
}

// Two-pass path: pass 2 summed e^2 * G over EVERY column; the columns that turned out to be
// background (a few per cent) are collected here, in index order, into a compact operand so that a
// short ragged pass can take their share out again.  gridDim.y blocks per instance: each repeats the
// (cheap) ordered compaction and copies its share of the rows.
__global__ void __launch_bounds__(1024)
tc_bg_gather_kernel(const float *__restrict__ fgcol, int Cn, int Np, int C, int B,
                    const __nv_bfloat16 *__restrict__ Bm /* (2, B*Np, C) */, const float4 *__restrict__ Gall,
                    __nv_bfloat16 *__restrict__ Bc /* (2, B*Np, C) */, float4 *__restrict__ Gc,
                    int *__restrict__ strip_count /* (B) */, int n_planes /* 2: hi + lo, 1: hi only */) {
    __shared__ int s_list[4096];
    __shared__ int s_warp[32];
    __shared__ int s_total;
    pdl_enter();
    const int b = blockIdx.x, tid = threadIdx.x, per = (Np + 1023) / 1024;  // Np <= 4096
    const int part_id = blockIdx.y, n_parts = gridDim.y;
    int flags = 0, cnt = 0;
    for (int u = 0; u < per; ++u) {
        const int j = tid * per + u;
        const int isbg = j < Cn && fgcol[(size_t)b * Np + j] == 0.f;
        flags |= isbg << u;
        cnt += isbg;
    }
    // block-wide exclusive scan of cnt
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, incl, o);
        if (lane_id() >= o) incl += v;
    }
    if (lane_id() == 31) s_warp[warp_id()] = incl;
    __syncthreads();
    if (warp_id() == 0) {
        int w = s_warp[lane_id()];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(kFull, w, o);
            if (lane_id() >= o) w += v;
        }
        s_warp[lane_id()] = w;
        if (lane_id() == 31) s_total = w;
    }
    __syncthreads();
    int pos = incl - cnt + (warp_id() ? s_warp[warp_id() - 1] : 0);
    for (int u = 0; u < per; ++u)
        if (flags >> u & 1) s_list[pos++] = tid * per + u;
    __syncthreads();
    const int nbg = s_total, padded = (nbg + TN - 1) / TN * TN;
    if (tid == 0 && part_id == 0) strip_count[b] = padded / TN;
    // gather the rows (hi and lo halves, 16 bytes per thread and trip), zero the padding
    const int v4 = C / 8;  // uint4 per row half
    const size_t half_rows = (size_t)B * Np;
    for (int e = part_id * 1024 + tid; e < padded * n_planes * v4; e += 1024 * n_parts) {
        const int q = e / (n_planes * v4), r = e - q * n_planes * v4, part = r / v4, x = r - part * v4;
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (q < nbg)
            val = __ldg(reinterpret_cast<const uint4 *>(Bm + (part * half_rows + (size_t)b * Np + s_list[q]) * C) + x);
        reinterpret_cast<uint4 *>(Bc + (part * half_rows + (size_t)b * Np + q) * C)[x] = val;
    }
    for (int q = part_id * 1024 + tid; q < padded; q += 1024 * n_parts)
        Gc[(size_t)b * Np + q] = q < nbg ? Gall[(size_t)b * Np + s_list[q]] : make_float4(0.f, 0.f, 0.f, 0.f);
}

// after pass 3: assignment_score and the soft correspondences (MU:324-330)
__global__ void __launch_bounds__(256)
tc_finish_wsum_kernel(const float *__restrict__ rowpart4, const float *__restrict__ inv_rowsum,
                      const float *__restrict__ w1, int R, int Mp, int n_part,
                      float *__restrict__ rowsumA, float *__restrict__ pred,
                      const float *__restrict__ corr4 /* (B,n_part_c,4,Mp) or null */, int n_part_c,
                      const int *__restrict__ strip_count /* (B): live strips of corr4 */,
                      int parts_per_strip) {
    // sub-thread k folds component k (weight, x, y, z) of row i over all partials
    __shared__ float red4[kFinishSub][kFinishItems];
    pdl_enter();
    const int b = blockIdx.y, item = threadIdx.x % kFinishItems, k = threadIdx.x / kFinishItems;
    const int i = blockIdx.x * kFinishItems + item;
    float ak = 0.f;
    if (i < R) {
        const float *src = rowpart4 + ((size_t)b * n_part * 4 + k) * Mp + i;
#pragma unroll 8
        for (int t = 0; t < n_part; ++t) ak += src[(size_t)t * 4 * Mp];
        if (corr4) {  // take the background columns' share out again (parts_per_strip partials per live strip)
            float c = 0.f;
            const int live = parts_per_strip * strip_count[b];
            const float *cs = corr4 + ((size_t)b * n_part_c * 4 + k) * Mp + i;
#pragma unroll 4
            for (int t = 0; t < live; ++t) c += cs[(size_t)t * 4 * Mp];
            ak -= c;
        }
    }
    red4[k][item] = ak;
    __syncthreads();
    if (k != 0 || i >= R) return;
    const float a[4] = {red4[0][item], red4[1][item], red4[2][item], red4[3][item]};
    const size_t ri = (size_t)b * R + i;
    const float f = w1[ri] * inv_rowsum[(size_t)b * Mp + i];
    const float sA = a[0] * f;
    rowsumA[ri] = sA;                       // assignment_score (MU:330)
    const float inv = 1.0f / (sA + 1e-6f);  // MU:327
    pred[ri * 3 + 0] = a[1] * f * inv;
    pred[ri * 3 + 1] = a[2] * f * inv;
    pred[ri * 3 + 2] = a[3] * f * inv;
}

// ---------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) !=
            cudaSuccess || qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
    return fn;
}

// 2D bf16 tensor (rows x C), box = 128 rows x 64 columns, SWIZZLE_128B
static int make_operand_map(CUtensorMap *map, const void *base, size_t rows, int C, int box_rows = BM) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return PEM_E_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? PEM_OK : PEM_E_UNSUPPORTED;
}

struct Ws {
    __nv_bfloat16 *A, *Bm;
    float *rowpart;  // (B, 4*n_strips, 4, Mp): the largest of the per-pass row partial layouts
    float *colpart;  // (2*n_msteps, B, Np): column partial planes (sums, then maxima)
    float *bg_a, *bg_b, *s00;  // background column / row of the cosine matrix and their corner
    float *lbg;                // (B,2) log2 of the background row / column sums
    // two-pass path: row maxima next to the weighted sums, background-column correction
    float *rowmax, *rowpart_c, *fgcol;
    float4 *Gall, *Gc;
    __nv_bfloat16 *Bc;
    int *strip_count;
    float *inv_rowsum, *lrs, *inv_colsum, *lcs, *w1, *w2, *rowsumA, *pred;
    float4 *G;
    int *counters;
    size_t bytes;
};

static int pad_to(int x, int g) { return (x + g - 1) / g * g; }

// rows = GEMM rows / columns per instance (without the background token)
static Ws carve(void *base, int B, int rows1, int rows2, int C) {
    const size_t Mp = pad_to(rows1, TM), Np = pad_to(rows2, TN), ns = Np / TN, nm = Mp / TM;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t o = off;
        off += (bytes + 1023) / 1024 * 1024;
        return (char *)base + o;
    };
    Ws w;
    w.A = (__nv_bfloat16 *)take(2 * B * Mp * C * 2);
    w.Bm = (__nv_bfloat16 *)take(2 * B * Np * C * 2);
    w.rowpart = (float *)take(B * 4 * ns * 4 * Mp * 4);  // up to 4 column parts per strip
    w.colpart = (float *)take(2 * nm * B * Np * 4);
    w.bg_a = (float *)take(B * Mp * 4);
    w.bg_b = (float *)take(B * Np * 4);
    w.s00 = (float *)take(B * 4);
    w.lbg = (float *)take(B * 8);
    w.rowmax = (float *)take(B * 4 * ns * Mp * 4);
    w.rowpart_c = (float *)take(B * 4 * ns * 4 * Mp * 4);
    w.fgcol = (float *)take(B * Np * 4);
    w.Gall = (float4 *)take(B * Np * 16);
    w.Gc = (float4 *)take(B * Np * 16);
    w.Bc = (__nv_bfloat16 *)take(2 * B * Np * C * 2);
    w.strip_count = (int *)take(B * 4);
    w.inv_rowsum = (float *)take(B * Mp * 4);
    w.lrs = (float *)take(B * Mp * 4);
    w.inv_colsum = (float *)take(B * Np * 4);
    w.lcs = (float *)take(B * Np * 4);
    w.w1 = (float *)take((size_t)B * rows1 * 4);
    w.w2 = (float *)take((size_t)B * rows2 * 4);
    w.rowsumA = (float *)take((size_t)B * rows1 * 4);
    w.pred = (float *)take((size_t)B * rows1 * 12);
    w.G = (float4 *)take(B * Np * 16);
    w.counters = (int *)take(B * 3 * 4);
    w.bytes = off;
    return w;
}

static int num_sms() { return pem_num_sms(); }

// Tuning knobs of the single-product mode.  Constants in the shipped library; a -DPEM_DEV_KNOBS build (make DEV=1,
// tools/pass_probe.py) makes them settable.
#ifdef PEM_DEV_KNOBS
static int g_epi_warps_single = kEpiWarpsSingle;  // epilogue warps of the single-product mode (8 or 16)
static int g_ring_single = kMaxStages;            // ring slots of the single-product mode
static int g_dev_steps_per_task = 0;              // row steps per task when the caller leaves the choice open
#else
static constexpr int g_epi_warps_single = kEpiWarpsSingle;
static constexpr int g_ring_single = kMaxStages;
#endif

// epilogue warps per CTA / row partials per strip for these parameters
static int epw_of(const Params &p) { return p.single ? g_epi_warps_single : 8; }

// rows1 x rows2 GEMM per instance
static int fill_common(Params &p, int B, int rows1, int rows2, int C, float temp, int single = 0, int steps_per_task = 0) {
#ifdef PEM_DEV_KNOBS
    if (steps_per_task == 0) steps_per_task = g_dev_steps_per_task;
#endif
    if (C % BK != 0 || C > BK * kMaxKC) return PEM_E_UNSUPPORTED;
    if (!(temp > 0.f) || 2.0f / temp > 80.f) return PEM_E_UNSUPPORTED;  // fixed-shift softmax range
    p.B = B; p.R = rows1; p.Cn = rows2;
    p.Mp = pad_to(rows1, TM); p.Np = pad_to(rows2, TN);
    p.n_msteps = p.Mp / TM; p.n_strips = p.Np / TN; p.KC = C / BK;
    // Row steps per task: short tasks balance the CTA pairs (a pair runs ceil(tasks / pairs) of
    // them) at the price of one strip reload (hidden) and one column fold per task.  With one
    // product per element the MMAs of a step take a third of the time and the pass runs at the
    // L2 -> shared-memory streaming rate (~16 B/clk per SM measured): a strip reload then costs about
    // as much as half a step of streamed rows, so long tasks win (statistics pass of 16 instances:
    // 46 us at 2 steps per task, 40 us at 8).
    p.single = single;
    const double reload = single ? 0.75 : 0.25;
    const int pairs = max(1, num_sms() / 2);
    int best = p.n_msteps;
    double best_cost = 1e30;
    for (int spt = p.n_msteps; spt >= 1; --spt) {
        const int splits = (p.n_msteps + spt - 1) / spt;
        const long tasks = (long)B * p.n_strips * splits;
        const long rounds = (tasks + pairs - 1) / pairs;
        const double cost = (double)rounds * (spt + reload);
        if (cost < best_cost) { best_cost = cost; best = spt; }
    }
    p.spt = steps_per_task > 0 ? min(steps_per_task, p.n_msteps) : best;
    p.spt_fixed = steps_per_task > 0;
    p.n_splits = (p.n_msteps + p.spt - 1) / p.spt;
    p.ring = g_ring_single;
    p.temp = temp;
    p.c1 = 1.4426950408889634f / temp;
    p.c2 = 1.4426950408889634f / temp;
    return PEM_OK;
}

// Parameters of the ragged background-column correction launch.  Only the first strip_count[b] strips of an
// instance are live there (one, as a rule): short tasks, so that the few live (instance, strip) pairs spread
// over the CTA pairs instead of 8 steps in a row on 16 of them.
static Params ragged_params(const Params &p, int B) {
    Params pc = p;
    const int pairs = max(1, num_sms() / 2);
    int best = 1;
    double best_cost = 1e30;
    for (int spt = p.n_msteps; spt >= 1; --spt) {
        const int splits = (p.n_msteps + spt - 1) / spt;
        const long rounds = ((long)B * splits + pairs - 1) / pairs;
        const double cost = (double)rounds * (spt + (p.single ? 0.75 : 0.25));
        if (cost <= best_cost) { best_cost = cost; best = spt; }
    }
    pc.spt = p.spt_fixed ? p.spt : best;
    pc.n_splits = (p.n_msteps + pc.spt - 1) / pc.spt;
    return pc;
}

// normalise + split both operands (and the background dots when skip = 1)
static void launch_prep(const Params &p, int B, int C, cudaStream_t st, const float *f1, __nv_bfloat16 *A,
                        const float *f2, __nv_bfloat16 *Bm, int normalize, int skip, float *bg_a, float *bg_b,
                        float *s00) {
    const int grid = min((B * (p.Mp + p.Np) / kPrepRows + 7) / 8, 3 * num_sms());  // persistent, 3 CTAs per SM
    auto go = [&](auto kern) {
        pem_launch_pdl(kern, dim3(grid), dim3(256), 0, st, f1, p.R, p.Mp, A, f2, p.Cn, p.Np, Bm, B, C, normalize, skip,
                       bg_a, bg_b, s00, p.single);
    };
    if (C <= 128) go(tc_prep_kernel<1>);
    else go(tc_prep_kernel<2>);
}

template <int MODE, int EPW, bool FAST = false>
static int launch_pass_w(const CUtensorMap &tmA, const CUtensorMap &tmB, const Params &p, cudaStream_t st) {
    constexpr int kThreads = 64 + 32 * EPW;
    const size_t smem = sizeof(Smem) + 1024;
    static PemSmemCache cached(0);
    if (int rc = pem_ensure_dyn_smem(fine_tc_kernel<MODE, EPW, FAST>, smem, &cached)) return rc;
    const int n_tasks = p.B * p.n_strips * p.n_splits;
    const int pairs = min(max(1, num_sms() / 2), n_tasks);  // one cluster of two CTAs per SM pair
    pem_launch_pdl(fine_tc_kernel<MODE, EPW, FAST>, dim3(2 * pairs), dim3(kThreads), smem, st, tmA, tmB, p);
    return pem_launch_status();
}
// the unrolled form of the kernel (see fine_tc_kernel): single product, C = 256, 16 epilogue warps; the
// role probes of the development build (16, 32, 64) keep the generic loops
static bool fast_form(const Params &p) {
    return p.single && p.KC == 4 && epw_of(p) == 16 && (p.dev & (16 | 32 | 64)) == 0;
}

template <int MODE>
static int launch_pass(const CUtensorMap &tmA, const CUtensorMap &tmB, const Params &p, cudaStream_t st) {
    if (MODE != kStore && epw_of(p) == 16) {
        // the unrolled producer / issuer (see fine_tc_kernel): single product, C = 256; probes keep the generic loops
        if (fast_form(p)) return launch_pass_w<MODE, MODE == kStore ? 8 : 16, MODE != kStore>(tmA, tmB, p, st);
        return launch_pass_w<MODE, MODE == kStore ? 8 : 16>(tmA, tmB, p, st);
    }
    return launch_pass_w<MODE, 8>(tmA, tmB, p, st);
}

}  // namespace tc
}  // namespace pem

using namespace pem;
using namespace pem::tc;

#ifdef PEM_DEV_KNOBS
extern "C" void pem_dev_set_fine_steps_per_task(int v) { g_dev_steps_per_task = v; }
extern "C" void pem_dev_set_fine_epi_warps(int v) { g_epi_warps_single = v == 8 ? 8 : 16; }
extern "C" void pem_dev_set_fine_ring(int v) { g_ring_single = v < 2 ? 2 : v > kMaxStages ? kMaxStages : v; }
static long long *g_dev_trace = nullptr;
extern "C" void pem_dev_set_fine_trace(long long *buf) { g_dev_trace = buf; }  // 4 roles x kTraceLen x (tag, clock)
#endif

// fine.cu
namespace pem {
int fine_tail_launch(const float *pred, const float *pts1, const float *rowsumA, const float *w1,
                     const float *model, int B, int N1, int Nm, float dis_thres, int *counters,
                     float *R, float *t, float *score, cudaStream_t st);
}

extern "C" size_t pem_fine_fused_workspace_bytes(int B, int N1, int N2, int C) {
    if (B <= 0 || N1 <= 0 || N2 <= 0 || C <= 0) return 0;
    return carve(nullptr, B, N1 + 1, N2 + 1, C).bytes;  // covers both entry points
}

// f1 (B,N1p,C), f2 (B,N2p,C) -> logits (B,N1p,N2p); every row takes part in the GEMM
extern "C" int pem_feature_similarity_tc(const float *f1, const float *f2, int B, int N1p, int N2p,
                                         int C, float temp, int normalize, float *atten,
                                         void *workspace, size_t workspace_bytes,
                                         pem_stream_t stream) {
    PEM_REQUIRE_PTR(f1);
    PEM_REQUIRE_PTR(f2);
    PEM_REQUIRE_PTR(atten);
    PEM_REQUIRE(B > 0 && N1p > 0 && N2p > 0 && C > 0, PEM_E_SHAPE);
    PEM_REQUIRE((uintptr_t)f1 % 16 == 0 && (uintptr_t)f2 % 16 == 0, PEM_E_UNSUPPORTED);  // 16-byte row loads
    Params p{};
    if (int rc = fill_common(p, B, N1p, N2p, C, temp)) return rc;
    PEM_REQUIRE(workspace != nullptr &&
                    workspace_bytes >= pem_fine_fused_workspace_bytes(B, N1p - 1, N2p - 1, C),
                PEM_E_WORKSPACE);
    cudaStream_t st = pem_cu_stream(stream);
    Ws w = carve(workspace, B, N1p, N2p, C);
    launch_prep(p, B, C, st, f1, w.A, f2, w.Bm, normalize, 0, nullptr, nullptr, nullptr);
    CUtensorMap tmA, tmB;
    if (int rc = make_operand_map(&tmA, w.A, (size_t)2 * B * p.Mp, C)) return rc;
    if (int rc = make_operand_map(&tmB, w.Bm, (size_t)2 * B * p.Np, C)) return rc;
    p.atten = atten;
    p.off = 0;
    return launch_pass<kStore>(tmA, tmB, p, st);
}

extern "C" int pem_fine_fused_ex(const float *f1, const float *f2, const float *pts1, const float *pts2,
                                 const float *model, int B, int N1, int N2, int C, int Nm, float temp,
                                 float dis_thres, int precision, float *R, float *t, float *score,
                                 float *w1_out, float *w2_out, float *rowsum_out, float *pred_out,
                                 void *workspace, size_t workspace_bytes, pem_stream_t stream) {
    // option bits ride on the precision word (pem_b200.h): PEM_FINE_THREE_PASS, PEM_FINE_STEPS_PER_TASK(n)
    const int options = precision & ~0xff;
    precision &= 0xff;
    PEM_REQUIRE((options & ~(PEM_FINE_THREE_PASS | PEM_FINE_STEPS_PER_TASK(0xff))) == 0, PEM_E_UNSUPPORTED);
    PEM_REQUIRE(precision == PEM_PREC_SPLIT_BF16 || precision == PEM_PREC_FP16, PEM_E_UNSUPPORTED);
    PEM_REQUIRE_PTR(f1);
    PEM_REQUIRE_PTR(f2);
    PEM_REQUIRE_PTR(pts1);
    PEM_REQUIRE_PTR(pts2);
    PEM_REQUIRE_PTR(model);
    PEM_REQUIRE_PTR(R);
    PEM_REQUIRE_PTR(t);
    PEM_REQUIRE_PTR(score);
    PEM_REQUIRE(B > 0 && N1 > 0 && N2 > 0 && C > 0 && Nm > 0 && B <= 65535, PEM_E_SHAPE);
    PEM_REQUIRE((uintptr_t)f1 % 16 == 0 && (uintptr_t)f2 % 16 == 0, PEM_E_UNSUPPORTED);  // 16-byte row loads
    Params p{};
    // the background token stays out of the GEMM
    if (int rc = fill_common(p, B, N1, N2, C, temp, precision == PEM_PREC_FP16, (options >> 16) & 0xff)) return rc;
    PEM_REQUIRE(workspace != nullptr &&
                    workspace_bytes >= pem_fine_fused_workspace_bytes(B, N1, N2, C),
                PEM_E_WORKSPACE);
    cudaStream_t st = pem_cu_stream(stream);
    Ws w = carve(workspace, B, N1, N2, C);
    float *w1 = w1_out ? w1_out : w.w1;
    float *rsA = rowsum_out ? rowsum_out : w.rowsumA;
    float *pred = pred_out ? pred_out : w.pred;

    launch_prep(p, B, C, st, f1, w.A, f2, w.Bm, 1, 1, w.bg_a, w.bg_b, w.s00);
    CUtensorMap tmA, tmB;
    if (int rc = make_operand_map(&tmA, w.A, (size_t)2 * B * p.Mp, C)) return rc;
    if (int rc = make_operand_map(&tmB, w.Bm, (size_t)2 * B * p.Np, C)) return rc;

    const int mx = max(p.Mp, p.Np);
    const int pps = epw_of(p) / 4;  // row partials per strip: one per epilogue warp of a quadrant
    const int n_part = pps * p.n_strips, planes = 2 * p.n_splits;
    const bool two_pass = !(options & PEM_FINE_THREE_PASS) && p.Np <= 4096;
    // pass 1: softmax statistics
    p.rowpart = w.rowpart;
    p.colpart = w.colpart;
    if (int rc = launch_pass<kStats>(tmA, tmB, p, st)) return rc;
    pem_launch_pdl(tc_finish_stats_kernel, dim3((mx + kFinishItems - 1) / kFinishItems + 1, B), dim3(256), 0, st, 
        w.rowpart, w.colpart, w.bg_a, w.bg_b, p.R, p.Cn, w.s00, p.Mp, p.Np, n_part, planes, p.c1, p.c2,
        w.inv_rowsum, w.lrs, w.inv_colsum, w.lcs, w.lbg, pts2, two_pass ? w.Gall : nullptr);
    p.lrs = w.lrs;
    p.lcs = w.lcs;
    if (two_pass) {
        // pass 2: foreground labels AND the weighted row sums over EVERY column (the column labels
        // are not known yet; ~96 % of the columns turn out foreground)
        p.rowmaxpart = w.rowmax;
        p.G = w.Gall;
        p.rowpart4 = w.rowpart;
        p.bg_b = w.bg_b;
        p.lbg = w.lbg;
        p.early = 1;  // the predecessor (tc_finish_stats_kernel) does not touch the operands
        if (int rc = launch_pass<kLabelW>(tmA, tmB, p, st)) return rc;
        p.early = 0;
        const int n_part_m = n_part;
        pem_launch_pdl(tc_finish_label_kernel, dim3((mx + kFinishItems - 1) / kFinishItems, B), dim3(256), 0, st, 
            w.rowmax, w.colpart, w.bg_a, w.bg_b, w.lbg, w.inv_colsum, pts2, p.R, p.Cn, p.Mp, p.Np, n_part_m,
            planes, p.c1, p.c2, w1, w2_out, nullptr, w.fgcol);
        // the background columns, compacted, and a short ragged pass that takes their share out again
        pem_launch_pdl(tc_bg_gather_kernel, dim3(B, 8), dim3(1024), 0, st, w.fgcol, p.Cn, p.Np, C, B, w.Bm, w.Gall, w.Bc, w.Gc,
                                                w.strip_count, p.single ? 1 : 2);
        CUtensorMap tmC;
        if (int rc = make_operand_map(&tmC, w.Bc, (size_t)2 * B * p.Np, C)) return rc;
        Params pc = ragged_params(p, B);
        pc.G = w.Gc;
        pc.rowpart4 = w.rowpart_c;
        pc.strip_count = w.strip_count;
        pc.strip_major = 1;
        if (int rc = launch_pass<kWsum>(tmA, tmC, pc, st)) return rc;
        pem_launch_pdl(tc_finish_wsum_kernel, dim3((N1 + kFinishItems - 1) / kFinishItems, B), dim3(256), 0, st, 
            w.rowpart, w.inv_rowsum, w1, p.R, p.Mp, n_part_m, rsA, pred, w.rowpart_c, n_part, w.strip_count, pps);
    } else {
        // pass 2: foreground labels
        p.rowmaxpart = w.rowpart;
        p.early = 1;  // the predecessors (finish kernels) do not touch the operands
        if (int rc = launch_pass<kLabel>(tmA, tmB, p, st)) return rc;
        pem_launch_pdl(tc_finish_label_kernel, dim3((mx + kFinishItems - 1) / kFinishItems, B), dim3(256), 0, st, 
            w.rowpart, w.colpart, w.bg_a, w.bg_b, w.lbg, w.inv_colsum, pts2, p.R, p.Cn, p.Mp, p.Np, n_part,
            planes, p.c1, p.c2, w1, w2_out, w.G, nullptr);
        // pass 3: weighted row sums -> correspondences
        p.G = w.G;
        p.rowpart4 = w.rowpart;
        if (int rc = launch_pass<kWsum>(tmA, tmB, p, st)) return rc;
        pem_launch_pdl(tc_finish_wsum_kernel, dim3((N1 + kFinishItems - 1) / kFinishItems, B), dim3(256), 0, st, 
            w.rowpart, w.inv_rowsum, w1, p.R, p.Mp, n_part, rsA, pred, nullptr, 0, nullptr, pps);
    }
    // weighted Procrustes + pose score (fine.cu)
    return fine_tail_launch(pred, pts1, rsA, w1, model, B, N1, Nm, dis_thres, w.counters, R, t,
                            score, st);
}

extern "C" int pem_fine_fused(const float *f1, const float *f2, const float *pts1, const float *pts2,
                              const float *model, int B, int N1, int N2, int C, int Nm, float temp,
                              float dis_thres, float *R, float *t, float *score, float *w1_out,
                              float *w2_out, float *rowsum_out, float *pred_out, void *workspace,
                              size_t workspace_bytes, pem_stream_t stream) {
    return pem_fine_fused_ex(f1, f2, pts1, pts2, model, B, N1, N2, C, Nm, temp, dis_thres, PEM_PREC_SPLIT_BF16, R,
                             t, score, w1_out, w2_out, rowsum_out, pred_out, workspace, workspace_bytes, stream);
}

// Development: time one GEMM pass (mode 1 stats, 2 label, 3 wsum) with parts of the kernel disabled
// (Params::dev).  Modes 2-3 read the constants a previous pem_fine_fused call left in `workspace`.
extern "C" int pem_dev_fine_pass(const float *f1, const float *f2, int B, int N1, int N2, int C, float temp,
                                 int flags, int mode, int prep, void *workspace, size_t workspace_bytes,
                                 pem_stream_t stream) {
    Params p{};
    if (int rc = fill_common(p, B, N1, N2, C, temp, (flags >> 3) & 1)) return rc;
    cudaStream_t st = pem_cu_stream(stream);
    Ws w = carve(workspace, B, N1, N2, C);
    if (workspace_bytes < w.bytes) return PEM_E_WORKSPACE;
    if (prep)
        launch_prep(p, B, C, st, f1, w.A, f2, w.Bm, 1, 1, w.bg_a, w.bg_b, w.s00);
    if (mode == 0) return pem_launch_status();  // the operand preparation only
    CUtensorMap tmA, tmB;
    if (int rc = make_operand_map(&tmA, w.A, (size_t)2 * B * p.Mp, C)) return rc;
    if (int rc = make_operand_map(&tmB, w.Bm, (size_t)2 * B * p.Np, C)) return rc;
#ifdef PEM_DEV_KNOBS
    p.trace = g_dev_trace;
#endif
    p.dev = flags & ~8;  // (bit 3: the single-product fp16 mode; `prep` must have run in the same mode)
    p.rowpart = p.rowmaxpart = p.rowpart4 = w.rowpart;
    p.colpart = w.colpart;
    p.lrs = w.lrs;
    p.lcs = w.lcs;
    p.G = w.G;
    if (mode == 4) {  // the merged pass of the two-pass path
        p.rowmaxpart = w.rowmax;
        p.G = w.Gall;
        p.bg_b = w.bg_b;
        p.lbg = w.lbg;
        return launch_pass<kLabelW>(tmA, tmB, p, st);
    }
    if (mode == 5) {  // the ragged background-column correction launch of the two-pass path
        CUtensorMap tmC;
        if (int rc = make_operand_map(&tmC, w.Bc, (size_t)2 * B * p.Np, C)) return rc;
        Params pc = ragged_params(p, B);
        pc.G = w.Gc;
        pc.rowpart4 = w.rowpart_c;
        pc.strip_count = w.strip_count;
        pc.strip_major = 1;
        return launch_pass<kWsum>(tmA, tmC, pc, st);
    }
    return mode == 3 ? launch_pass<kWsum>(tmA, tmB, p, st)
                     : mode == 2 ? launch_pass<kLabel>(tmA, tmB, p, st) : launch_pass<kStats>(tmA, tmB, p, st);
}

// One reducing GEMM pass on its own (profiling hook, see pem_b200.h)
extern "C" int pem_fine_fused_pass_ex(const float *f1, const float *f2, int B, int N1, int N2, int C, float temp,
                                      int pass, int precision, void *workspace, size_t workspace_bytes,
                                      pem_stream_t stream) {
    PEM_REQUIRE(precision == PEM_PREC_SPLIT_BF16 || precision == PEM_PREC_FP16, PEM_E_UNSUPPORTED);
    PEM_REQUIRE_PTR(f1);
    PEM_REQUIRE_PTR(f2);
    PEM_REQUIRE(B > 0 && N1 > 0 && N2 > 0 && C > 0, PEM_E_SHAPE);
    PEM_REQUIRE(pass >= 0 && pass <= 5, PEM_E_UNSUPPORTED);
    PEM_REQUIRE(workspace != nullptr && workspace_bytes >= pem_fine_fused_workspace_bytes(B, N1, N2, C),
                PEM_E_WORKSPACE);
    // pass 0: the operand preparation alone (what precedes the passes in a step and leaves the
    // operands in L2 for them)
    return pem_dev_fine_pass(f1, f2, B, N1, N2, C, temp, precision == PEM_PREC_FP16 ? 8 : 0, pass, pass == 0,
                             workspace, workspace_bytes, stream);
}

extern "C" int pem_fine_fused_pass(const float *f1, const float *f2, int B, int N1, int N2, int C, float temp,
                                   int pass, void *workspace, size_t workspace_bytes, pem_stream_t stream) {
    return pem_fine_fused_pass_ex(f1, f2, B, N1, N2, C, temp, pass, PEM_PREC_SPLIT_BF16, workspace, workspace_bytes,
                                  stream);
}
