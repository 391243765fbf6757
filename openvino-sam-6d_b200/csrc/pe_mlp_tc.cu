// PositionalEncoding's per-scale branch, fused (model/fine_point_matching.py:127-141 of the reference;
// SURVEY.md section 8f, rank 1):
//
//   QueryAndGroup (neighbour - centre | neighbour)  ->  SharedMLP [6, 32, 64, 128] (three times 1x1
//   conv + BatchNorm (eval) + ReLU)  ->  amax over the neighbours
//
// The reference writes the (B,6,N,ns) grouped tensor and the (B,32|64|128,N,ns) activations to memory
// (ns = 32 / 64: 0.8 / 1.6 GB per step at 32 clouds of 2048 points for the last layer alone) only to
// max-pool them.  Here a row = one (centre, neighbour) pair; 128 rows at a time go through the three
// layers on the tensor cores without leaving the SM:
//
//   * row threads (warps 1-4, thread = row = TMEM lane) build the 6 inputs of their row from the
//     neighbour lists and the cloud in shared memory, as bf16 hi + lo, in a SWIZZLE_128B operand tile;
//   * warp 0 issues tcgen05.mma kind::f16 M128 x N{32,64,128} x K16: A_hi W_hi + A_hi W_lo + A_lo W_hi
//     (split-bf16, fp32-class accuracy) against weights that stay resident in shared memory;
//   * the row threads read the layer's accumulator out of TMEM, add the BatchNorm shift (its scale is
//     folded into the weights by the caller), apply the ReLU, split the result again and write it as the next layer's operand tile;
//   * after the third layer a 31-shuffle transpose reduce takes the maximum over the 32 rows of a warp
//     (one centre at ns = 32, half a centre at ns = 64) and lane L writes channel L of each 32-channel
//     group: (B, N, ldo) point-major, ready for the final Linear / Conv1d(256, out_dim).
// The epilogues dominate (224 accumulator columns per row through affine / ReLU / re-split), so for
// clouds of up to 2048 points a CTA runs TWO row groups on alternate tiles, each with its own operand
// tiles, TMEM columns and barriers: the MMAs of one group's layer overlap the other group's epilogue,
// and every scheduler has two warps to switch between.
#include "tc_ptx.cuh"

namespace pem {
namespace tc {

constexpr int kPmRows = 128;
constexpr int kPmC0 = 6, kPmC1 = 32, kPmC2 = 64, kPmC3 = 128;
constexpr int kPmMaxN = 4096;             // one row group; two groups: 2048 (the cloud shares shared memory)
constexpr int kPmD1 = 0, kPmD2 = 32, kPmD3 = 96, kPmDcols = 256;  // D1 | D2 | D3 = 32 + 64 + 128 columns per group

template <int G>
struct PmSmem {
    uint8_t w1[2][kPmC1 * 128];          // resident weights, rows = output channels, hi / lo
    uint8_t w2[2][kPmC2 * 128];
    uint8_t w3[2][kPmC3 * 128];
    uint8_t a[G][2][2][kPmRows * 128];   // per group: two operand tiles (ping-pong between layers), hi / lo
    float4 cloud[kPmMaxN / G];
    alignas(16) float shift[kPmC1 + kPmC2 + kPmC3];  // BatchNorm shift per channel (its scale is folded into the weights)
    float comb[G][4][kPmC3];             // per-quadrant maxima (ns = 64: two quadrants per centre)
    uint64_t bar_a[G], bar_d[G];
    uint32_t tmem_base;
};

// barrier 1 + g: the 128 row threads of group g; barrier 3: all row threads
__device__ __forceinline__ void pm_group_barrier(int g) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(kPmRows) : "memory");
}
template <int THREADS>
__device__ __forceinline__ void pm_rows_barrier() { asm volatile("bar.sync 3, %0;" ::"n"(THREADS) : "memory"); }

// 8 fp32 -> 8 bf16 hi and 8 bf16 lo, one 16-byte chunk each
__device__ __forceinline__ void pm_split8(const float (&v)[8], uint4 &hi, uint4 &lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
        const float2 hf = __bfloat1622float2(hh);
        const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * k] - hf.x, v[2 * k + 1] - hf.y);
        h[k] = *reinterpret_cast<const uint32_t *>(&hh);
        l[k] = *reinterpret_cast<const uint32_t *>(&ll);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// weights (Co, Ci) fp32 row-major -> K-major SWIZZLE_128B rows, hi / lo, zero padded to `kpad` inputs
__device__ __forceinline__ void pm_stage_weights(const float *__restrict__ w, int Co, int Ci, int kpad,
                                                 uint8_t *hi, uint8_t *lo) {
    for (int e = threadIdx.x; e < Co * (kpad / 8); e += blockDim.x) {
        const int r = e / (kpad / 8), c = e - r * (kpad / 8);
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (c * 8 + k < Ci) ? __ldg(w + (size_t)r * Ci + c * 8 + k) : 0.f;
        uint4 h, l;
        pm_split8(v, h, l);
        *reinterpret_cast<uint4 *>(hi + sw128_offset(r, c)) = h;
        *reinterpret_cast<uint4 *>(lo + sw128_offset(r, c)) = l;
    }
}

template <int NS, int G>
__global__ void __launch_bounds__(32 + G * kPmRows, 1)
pe_mlp_kernel(const float *__restrict__ xyz, const int32_t *__restrict__ idx, int B, int N, float eps,
              const float *__restrict__ w1, const float *__restrict__ w2, const float *__restrict__ w3,
              const float *__restrict__ shift, float *__restrict__ out, int ldo, int ch_off) {
    static_assert(kPmRows % NS == 0 || NS == 128, "whole centres per warp group");
    static_assert(NS == 32 || NS == 64 || NS == 128, "neighbours per centre");
    constexpr int kThreads = 32 + G * kPmRows;  // warp 0: MMA issuer; then 4 row warps per group
    constexpr uint32_t kTmemCols = G * kPmDcols;
    extern __shared__ uint8_t smem_raw[];
    PmSmem<G> &sm = *reinterpret_cast<PmSmem<G> *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = warp_id(), lane = lane_id(), tid = threadIdx.x;
    const int tpc = (N * NS + kPmRows - 1) / kPmRows;  // tiles per cloud (even when G = 2: host)
    // a CTA owns a contiguous range of SLOTS of G consecutive tiles; group g takes tile G*slot + g
    const int n_slots = B * tpc / G;
    const int s_begin = (int)((long long)n_slots * blockIdx.x / gridDim.x);
    const int s_end = (int)((long long)n_slots * (blockIdx.x + 1) / gridDim.x);

    if (tid == 0) {
        for (int g = 0; g < G; ++g) {
            mbar_init(smem_u32(&sm.bar_a[g]), kPmRows);
            mbar_init(smem_u32(&sm.bar_d[g]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&sm.tmem_base)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pm_stage_weights(w1, kPmC1, kPmC0, 16, sm.w1[0], sm.w1[1]);
    pm_stage_weights(w2, kPmC2, kPmC1, 32, sm.w2[0], sm.w2[1]);
    pm_stage_weights(w3, kPmC3, kPmC2, 64, sm.w3[0], sm.w3[1]);
    for (int c = tid; c < kPmC1 + kPmC2 + kPmC3; c += kThreads) sm.shift[c] = __ldg(shift + c);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;

    if (warp == 0) {
        // ================================ MMA issuer ==================================
        // per slot: layer 1 of every group, layer 2 of every group, layer 3 of every group -- while one
        // group's row threads turn an accumulator into the next operand, the other group's MMAs run
        constexpr uint32_t kI1 = make_idesc(kPmRows, kPmC1), kI2 = make_idesc(kPmRows, kPmC2),
                           kI3 = make_idesc(kPmRows, kPmC3);
        uint32_t round[G];  // one per layer and group: bar_a[g] completes once per round
        for (int g = 0; g < G; ++g) round[g] = 0;
        auto layer = [&](int g, int buf, const uint8_t *wh, const uint8_t *wl, int ksteps, uint32_t dcol,
                         uint32_t idesc) {
            mbar_wait(smem_u32(&sm.bar_a[g]), round[g] & 1);
            ++round[g];
            tc_fence_after();
            if (elect_one()) {
                const uint64_t ah = umma_desc_sw128(smem_u32(sm.a[g][buf][0])),
                               al = umma_desc_sw128(smem_u32(sm.a[g][buf][1]));
                const uint64_t bh = umma_desc_sw128(smem_u32(wh)), bl = umma_desc_sw128(smem_u32(wl));
                const uint32_t d = tmem_base + (uint32_t)(g * kPmDcols) + dcol;
                for (int ks = 0; ks < ksteps; ++ks) {  // +32 bytes per K = 16 step: +2 in the address field
                    tc_mma_bf16(d, ah + 2 * ks, bh + 2 * ks, idesc, ks != 0);
                    tc_mma_bf16(d, ah + 2 * ks, bl + 2 * ks, idesc, 1);
                    tc_mma_bf16(d, al + 2 * ks, bh + 2 * ks, idesc, 1);
                }
                tc_commit(smem_u32(&sm.bar_d[g]));
            }
            __syncwarp();
        };
        for (int slot = s_begin; slot < s_end; ++slot) {
#pragma unroll
            for (int g = 0; g < G; ++g) layer(g, 0, sm.w1[0], sm.w1[1], 1, kPmD1, kI1);
#pragma unroll
            for (int g = 0; g < G; ++g) layer(g, 1, sm.w2[0], sm.w2[1], 2, kPmD2, kI2);
#pragma unroll
            for (int g = 0; g < G; ++g) layer(g, 0, sm.w3[0], sm.w3[1], 4, kPmD3, kI3);
        }
    } else {
        // ================================ row threads =================================
        const int g = (warp - 1) >> 2;    // row group
        const int quad = warp & 3;        // TMEM lane quadrant this warp may read
        const int row = quad * 32 + lane; // tile row = TMEM lane
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(g * kPmDcols);
        uint32_t dround = 0;              // bar_d[g] completes once per layer
        int cloud_loaded = -1;
        auto wait_d = [&]() {
            mbar_wait(smem_u32(&sm.bar_d[g]), dround & 1);
            ++dround;
            tc_fence_after();
        };
        // accumulator columns [dcol, dcol + n) of this row -> BatchNorm affine + ReLU -> next operand tile
        auto relayer = [&](uint32_t dcol, int n, int ch0, int buf) {
            for (int c = 0; c < n; c += 32) {
                float v[32];
                tmem_ld32(lane_addr + dcol + c, v);
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                    const float4 s0 = lds_f4(&sm.shift[ch0 + c + g8 * 8]), s1 = lds_f4(&sm.shift[ch0 + c + g8 * 8 + 4]);
                    const float sh[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                    float y[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) y[k] = fmaxf(v[g8 * 8 + k] + sh[k], 0.f);
                    uint4 h, l;
                    pm_split8(y, h, l);
                    const uint32_t off = sw128_offset(row, (c >> 3) + g8);
                    *reinterpret_cast<uint4 *>(sm.a[g][buf][0] + off) = h;
                    *reinterpret_cast<uint4 *>(sm.a[g][buf][1] + off) = l;
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();
            mbar_arrive(smem_u32(&sm.bar_a[g]));
        };
        for (int slot = s_begin; slot < s_end; ++slot) {
            const int tile = slot * G + g;
            const int b = tile / tpc, r0 = (tile - b * tpc) * kPmRows;  // both tiles of a slot: same cloud
            if (b != cloud_loaded) {  // stage the cloud (every row thread is done with the previous one)
                pm_rows_barrier<G * kPmRows>();
                const float *pb = xyz + (size_t)b * N * 3;
                for (int k = tid - 32; k < N; k += G * kPmRows)
                    sm.cloud[k] = make_float4(__ldg(pb + 3 * k), __ldg(pb + 3 * k + 1), __ldg(pb + 3 * k + 2), 0.f);
                pm_rows_barrier<G * kPmRows>();
                cloud_loaded = b;
            }
            // ---- layer-1 operand: (neighbour - centre | neighbour) of this row
            const int grow = r0 + row, j = grow / NS;  // centre of this row
            const bool valid = j < N;
            {
                float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (valid) {
                    const int a = __ldg(idx + ((size_t)b * N) * NS + grow);
                    const float4 pa = sm.cloud[a], pj = sm.cloud[j];
                    x[0] = __fsub_rn(pa.x, __fadd_rn(pj.x, eps));
                    x[1] = __fsub_rn(pa.y, __fadd_rn(pj.y, eps));
                    x[2] = __fsub_rn(pa.z, __fadd_rn(pj.z, eps));
                    x[3] = pa.x; x[4] = pa.y; x[5] = pa.z;
                }
                uint4 h, l;
                pm_split8(x, h, l);
                const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4 *>(sm.a[g][0][0] + sw128_offset(row, 0)) = h;
                *reinterpret_cast<uint4 *>(sm.a[g][0][1] + sw128_offset(row, 0)) = l;
                *reinterpret_cast<uint4 *>(sm.a[g][0][0] + sw128_offset(row, 1)) = z;  // K = 16: inputs 8..15 are zero
                *reinterpret_cast<uint4 *>(sm.a[g][0][1] + sw128_offset(row, 1)) = z;
                fence_proxy_async_smem();
                mbar_arrive(smem_u32(&sm.bar_a[g]));
            }
            wait_d();
            relayer(kPmD1, kPmC1, 0, 1);               // -> layer-2 operand in tile 1
            wait_d();
            relayer(kPmD2, kPmC2, kPmC1, 0);           // -> layer-3 operand in tile 0
            wait_d();
            // ---- layer 3: affine + ReLU, maximum over the rows of each centre
            const int centre = (r0 + quad * 32) / NS;  // centre of this warp's 32 rows
#pragma unroll 1
            for (int c = 0; c < kPmC3; c += 32) {
                float v[32];
                tmem_ld32(lane_addr + kPmD3 + c, v);
#pragma unroll
                for (int k = 0; k < 32; k += 4) {
                    const float4 s4 = lds_f4(&sm.shift[kPmC1 + kPmC2 + c + k]);
                    v[k] = valid ? fmaxf(v[k] + s4.x, 0.f) : 0.f;
                    v[k + 1] = valid ? fmaxf(v[k + 1] + s4.y, 0.f) : 0.f;
                    v[k + 2] = valid ? fmaxf(v[k + 2] + s4.z, 0.f) : 0.f;
                    v[k + 3] = valid ? fmaxf(v[k + 3] + s4.w, 0.f) : 0.f;
                }
                const float m = warp_transpose_reduce<true>(v);  // lane L: channel c + L over 32 rows
                if (NS == 32) {
                    if (centre < N) out[((size_t)b * N + centre) * ldo + ch_off + c + lane] = m;
                } else {
                    sm.comb[g][quad][c + lane] = m;
                }
            }
            tc_fence_before();
            if (NS != 32) {  // combine the warps that share a centre
                constexpr int WPC = NS / 32;  // warps per centre: 2 or 4
                pm_group_barrier(g);
                if (quad % WPC == 0 && centre < N) {
                    for (int c = lane; c < kPmC3; c += 32) {
                        float m = sm.comb[g][quad][c];
#pragma unroll
                        for (int q = 1; q < WPC; ++q) m = fmaxf(m, sm.comb[g][quad + q][c]);
                        out[((size_t)b * N + centre) * ldo + ch_off + c] = m;
                    }
                }
                pm_group_barrier(g);  // comb is rewritten by the next tile
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols)
                     : "memory");
    }
}

}  // namespace tc
}  // namespace pem

using namespace pem;
using namespace pem::tc;

extern "C" int pem_pe_mlp(const float *xyz, const int32_t *idx, int B, int N, int nsample, float centre_eps,
                          const float *w1, const float *w2, const float *w3, const float *shift,
                          float *out, int ldo, int ch_off, pem_stream_t stream) {
    PEM_REQUIRE_PTR(xyz);
    PEM_REQUIRE_PTR(idx);
    PEM_REQUIRE_PTR(w1);
    PEM_REQUIRE_PTR(w2);
    PEM_REQUIRE_PTR(w3);
    PEM_REQUIRE_PTR(shift);
    PEM_REQUIRE_PTR(out);
    PEM_REQUIRE(B > 0 && N > 0 && ldo >= ch_off + kPmC3 && ch_off >= 0, PEM_E_SHAPE);
    PEM_REQUIRE(N <= kPmMaxN && (nsample == 32 || nsample == 64 || nsample == 128), PEM_E_UNSUPPORTED);
    const int sms = pem_num_sms();
    const long long tpc = ((long long)N * nsample + kPmRows - 1) / kPmRows;
    const bool two = N <= kPmMaxN / 2 && tpc % 2 == 0;  // two row groups per CTA on alternate tiles
    const long long n_slots = (long long)B * tpc / (two ? 2 : 1);
    const int grid = (int)(n_slots < sms ? n_slots : sms);
    cudaStream_t st = pem_cu_stream(stream);
    auto go = [&](auto kern, size_t smem, int threads, PemSmemCache *cached) -> int {
        if (int rc = pem_ensure_dyn_smem(kern, smem, cached)) return rc;
        kern<<<grid, threads, smem, st>>>(xyz, idx, B, N, centre_eps, w1, w2, w3, shift, out, ldo, ch_off);
        return pem_launch_status();
    };
    static PemSmemCache cached[6];
    const size_t s1 = sizeof(PmSmem<1>) + 1024, s2 = sizeof(PmSmem<2>) + 1024;
    const int t1 = 32 + kPmRows, t2 = 32 + 2 * kPmRows;
    if (two) {
        if (nsample == 32) return go(pe_mlp_kernel<32, 2>, s2, t2, &cached[0]);
        if (nsample == 64) return go(pe_mlp_kernel<64, 2>, s2, t2, &cached[1]);
        return go(pe_mlp_kernel<128, 2>, s2, t2, &cached[2]);
    }
    if (nsample == 32) return go(pe_mlp_kernel<32, 1>, s1, t1, &cached[3]);
    if (nsample == 64) return go(pe_mlp_kernel<64, 1>, s1, t1, &cached[4]);
    return go(pe_mlp_kernel<128, 1>, s1, t1, &cached[5]);
}
