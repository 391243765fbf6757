// Coarse pose-hypothesis stage of the PEM hot path: compute_coarse_Rt
// (utils/model_utils.py:204-275 of the reference, abbreviated MU below) split into the
// kernels the C ABI exposes:
//
//   coarse_assign      MU:230-242  dual softmax, fg masks, score**1.5, fp64 scan, normalise
//   coarse_hypotheses  MU:248-261  inverse-CDF triplets, 3-point Procrustes, residual
//   select_smallest    MU:262      top-P2 smallest residuals (radix select, as a set)
//   score_hypotheses   MU:267-271  transform + nearest model point + weighted score
//   select_best        MU:272-274  argmax, copy the winning (R,t)
//
// B200 notes: the 197x197 assignment matrix lives in shared memory for the whole of
// coarse_assign (155 KB of the 227 KB an SM offers); hypothesis solving keeps each 3x3
// problem in the registers of one thread (a warp = a tile of 32 hypotheses); scoring is the
// FP32-pipe-bound kernel: model points are staged once per CTA as (-2m, |m|^2) float4 so a
// pair costs 3 FFMA + 1 FMNMX, each LDS.128 broadcast is reused by up to 8 register-held
// query points per thread, and background points (w1 = 0) are compacted away up front.
#include <cooperative_groups.h>

#include "pem_common.cuh"
#include "svd3.cuh"

namespace cg = cooperative_groups;

namespace pem {

// ------------------------------------------------------------------------------------
// coarse_assign
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float pow15(float x) { return x * sqrtf(x); }

__global__ void __launch_bounds__(1024, 1)
coarse_assign_kernel(const float *__restrict__ atten, int N1, int N2, float *__restrict__ w1,
                     float *__restrict__ w2, float *__restrict__ cdf) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int R = N1 + 1, Cn = N2 + 1, tid = threadIdx.x, b = blockIdx.x;
    double *s_scan = reinterpret_cast<double *>(s_raw);  // 33 doubles
    float *S = reinterpret_cast<float *>(s_scan + 34);   // R*Cn
    float *rmax = S + R * Cn, *rsum = rmax + R, *cmax = rsum + R, *csum = cmax + Cn;
    float *w1s = csum + Cn, *w2s = w1s + R;              // indexed by full row / column index

    const float *A = atten + (size_t)b * R * Cn;
    for (int e = tid; e < R * Cn; e += 1024) S[e] = __ldg(A + e);
    __syncthreads();

    // Range of the logits.  When it is small (cosine similarity / temp always is: |s| <= 1/temp)
    // a single shift serves every row and column, and one exp per element replaces four.
    float lmax = -INFINITY, lmin = INFINITY;
    for (int e = tid; e < R * Cn; e += 1024) { lmax = fmaxf(lmax, S[e]); lmin = fminf(lmin, S[e]); }
    lmax = warp_max(lmax);
    lmin = warp_min(lmin);
    float *s_f = reinterpret_cast<float *>(s_scan);  // 68 floats of scratch, free until the scan
    if (lane_id() == 0) { s_f[warp_id()] = lmax; s_f[32 + warp_id()] = lmin; }
    __syncthreads();
    float gmax = s_f[lane_id()], gmin = s_f[32 + lane_id()];
    gmax = warp_max(gmax);
    gmin = warp_min(gmin);
    const bool fast = (gmax - gmin) < 80.0f;  // block-uniform
    __syncthreads();

    if (fast) {
        for (int e = tid; e < R * Cn; e += 1024) S[e] = expf(S[e] - gmax);
        __syncthreads();
        for (int i = warp_id(); i < R; i += 32) {
            float s = 0.f;
            for (int j = lane_id(); j < Cn; j += 32) s += S[i * Cn + j];
            s = warp_sum(s);
            if (lane_id() == 0) rsum[i] = 1.0f / s;
        }
        for (int j = warp_id(); j < Cn; j += 32) {  // odd row stride: conflict-free column walks
            float s = 0.f;
            for (int i = lane_id(); i < R; i += 32) s += S[i * Cn + j];
            s = warp_sum(s);
            if (lane_id() == 0) csum[j] = 1.0f / s;
        }
        __syncthreads();
        // P = softmax(dim=2) * softmax(dim=1) (MU:230)
        for (int e = tid; e < R * Cn; e += 1024) {
            const int i = e / Cn, j = e - i * Cn;
            const float v = S[e];
            S[e] = (v * rsum[i]) * (v * csum[j]);
        }
    } else {
        // general logits: per-row and per-column maxima (warp per row / per column)
        for (int i = warp_id(); i < R; i += 32) {
            float m = -INFINITY;
            for (int j = lane_id(); j < Cn; j += 32) m = fmaxf(m, S[i * Cn + j]);
            m = warp_max(m);
            float s = 0.f;
            for (int j = lane_id(); j < Cn; j += 32) s += expf(S[i * Cn + j] - m);
            s = warp_sum(s);
            if (lane_id() == 0) { rmax[i] = m; rsum[i] = s; }
        }
        for (int j = warp_id(); j < Cn; j += 32) {
            float m = -INFINITY;
            for (int i = lane_id(); i < R; i += 32) m = fmaxf(m, S[i * Cn + j]);
            m = warp_max(m);
            float s = 0.f;
            for (int i = lane_id(); i < R; i += 32) s += expf(S[i * Cn + j] - m);
            s = warp_sum(s);
            if (lane_id() == 0) { cmax[j] = m; csum[j] = s; }
        }
        __syncthreads();
        for (int e = tid; e < R * Cn; e += 1024) {
            const int i = e / Cn, j = e - i * Cn;
            const float v = S[e];
            S[e] = __fdiv_rn(expf(v - rmax[i]), rsum[i]) * __fdiv_rn(expf(v - cmax[j]), csum[j]);
        }
    }
    __syncthreads();

    // foreground masks (MU:231-234): the first maximum of the full row (column) is not index 0
    for (int i = 1 + warp_id(); i < R; i += 32) {
        float m = -INFINITY;
        for (int j = 1 + lane_id(); j < Cn; j += 32) m = fmaxf(m, S[i * Cn + j]);
        m = warp_max(m);
        if (lane_id() == 0) {
            const float f = (m > S[i * Cn]) ? 1.f : 0.f;
            w1s[i] = f;
            w1[(size_t)b * N1 + i - 1] = f;
        }
    }
    for (int j = 1 + warp_id(); j < Cn; j += 32) {
        float m = -INFINITY;
        for (int i = 1 + lane_id(); i < R; i += 32) m = fmaxf(m, S[i * Cn + j]);
        m = warp_max(m);
        if (lane_id() == 0) {
            const float f = (m > S[j]) ? 1.f : 0.f;
            w2s[j] = f;
            w2[(size_t)b * N2 + j - 1] = f;
        }
    }
    __syncthreads();

    // score**1.5 -> prefix sum in fp64 rounded to fp32 per element (torch CPU cumsum keeps a
    // double accumulator) -> divide by (last + 1e-8) in fp32 (MU:236-242)
    const int E = N1 * N2, chunk = (E + 1023) / 1024;
    const int e0 = min(E, tid * chunk), e1 = min(E, e0 + chunk);
    auto q_at = [&](int i, int j) {  // i,j are 0-based in the [1:,1:] block
        return pow15(S[(i + 1) * Cn + j + 1] * w1s[i + 1] * w2s[j + 1]);
    };
    double local = 0.0;
    {
        int i = e0 / N2, j = e0 - i * N2;
        for (int e = e0; e < e1; ++e) {
            local += (double)q_at(i, j);
            if (++j == N2) { j = 0; ++i; }
        }
    }
    // block exclusive scan of the 1024 partial sums
    double inc = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double n = __shfl_up_sync(kFull, inc, o);
        if (lane_id() >= o) inc += n;
    }
    if (lane_id() == 31) s_scan[warp_id()] = inc;
    __syncthreads();
    if (warp_id() == 0) {
        double v = s_scan[lane_id()];
        double winc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double n = __shfl_up_sync(kFull, winc, o);
            if (lane_id() >= o) winc += n;
        }
        s_scan[lane_id()] = winc - v;  // exclusive warp offsets
        if (lane_id() == 31) s_scan[32] = winc;
    }
    __syncthreads();
    double run = s_scan[warp_id()] + (inc - local);
    const float denom = (float)s_scan[32] + 1e-8f;
    float *out = cdf + (size_t)b * E;
    {
        int i = e0 / N2, j = e0 - i * N2;
        for (int e = e0; e < e1; ++e) {
            run += (double)q_at(i, j);
            out[e] = __fdiv_rn((float)run, denom);
            if (++j == N2) { j = 0; ++i; }
        }
    }
}

// Cluster version: the (N1+1) x (N2+1) matrix is split into kAsgCl row slabs, one per CTA of a
// thread-block cluster (one cluster per instance), so the front half of compute_coarse_Rt runs on
// kAsgCl SMs instead of one.  Row statistics are slab-local; column statistics, the column
// foreground test and the prefix-sum offsets are exchanged through distributed shared memory
// (four exchanges, each a burst of remote stores + one cluster barrier).  Arithmetic is the
// reference's: softmax factors exp(s - max)/sum per row and per column, fp64 scan rounded to fp32.
constexpr int kAsgCl = 4;

__global__ void __launch_bounds__(1024, 1)
coarse_assign_cluster_kernel(const float *__restrict__ atten, int N1, int N2, float *__restrict__ w1,
                             float *__restrict__ w2, float *__restrict__ cdf) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int b = blockIdx.x / kAsgCl;
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int R = N1 + 1, Cn = N2 + 1, tid = threadIdx.x;
    const int per = (R + kAsgCl - 1) / kAsgCl;
    const int r0 = rank * per, nr = max(0, min(per, R - r0));  // this CTA owns rows [r0, r0+nr)
    double *s_scan = reinterpret_cast<double *>(s_raw);          // 34 doubles
    double *s_tot = s_scan + 34;                                 // kAsgCl slab totals
    float *S = reinterpret_cast<float *>(s_tot + kAsgCl);        // per*Cn
    float *xch = S + per * Cn;                                   // [kAsgCl][Cn] exchange buffer A
    float *xch2 = xch + kAsgCl * Cn;                             // [kAsgCl][Cn] exchange buffer B
    float *rmax = xch2 + kAsgCl * Cn, *rinv = rmax + per;        // per each
    float *cmax = rinv + per, *cinv = cmax + Cn, *p0 = cinv + Cn, *w2s = p0 + Cn;  // Cn each
    float *w1s = w2s + Cn;                                       // per

    const float *A = atten + ((size_t)b * R + r0) * Cn;
    for (int e = tid; e < nr * Cn; e += 1024) S[e] = __ldg(A + e);
    __syncthreads();

    // ---- maxima: rows local, columns exchanged -------------------------------------------
    for (int i = warp_id(); i < nr; i += 32) {
        float m = -INFINITY;
        for (int j = lane_id(); j < Cn; j += 32) m = fmaxf(m, S[i * Cn + j]);
        m = warp_max(m);
        if (lane_id() == 0) rmax[i] = m;
    }
    for (int j = warp_id(); j < Cn; j += 32) {
        float m = -INFINITY;
        for (int i = lane_id(); i < nr; i += 32) m = fmaxf(m, S[i * Cn + j]);
        m = warp_max(m);
        if (lane_id() < kAsgCl) *cluster.map_shared_rank(&xch[rank * Cn + j], lane_id()) = m;
    }
    cluster.sync();
    for (int j = tid; j < Cn; j += 1024) {
        float m = xch[j];
#pragma unroll
        for (int q = 1; q < kAsgCl; ++q) m = fmaxf(m, xch[q * Cn + j]);
        cmax[j] = m;
    }
    __syncthreads();
    // ---- sums of exp: rows local, columns exchanged (fixed rank order: deterministic) -------
    for (int i = warp_id(); i < nr; i += 32) {
        float sm = 0.f;
        for (int j = lane_id(); j < Cn; j += 32) sm += expf(S[i * Cn + j] - rmax[i]);
        sm = warp_sum(sm);
        if (lane_id() == 0) rinv[i] = sm;
    }
    for (int j = warp_id(); j < Cn; j += 32) {
        float sm = 0.f;
        for (int i = lane_id(); i < nr; i += 32) sm += expf(S[i * Cn + j] - cmax[j]);
        sm = warp_sum(sm);
        if (lane_id() < kAsgCl) *cluster.map_shared_rank(&xch2[rank * Cn + j], lane_id()) = sm;
    }
    cluster.sync();
    for (int j = tid; j < Cn; j += 1024) {
        float sm = xch2[j];
#pragma unroll
        for (int q = 1; q < kAsgCl; ++q) sm += xch2[q * Cn + j];
        cinv[j] = sm;
    }
    __syncthreads();
    // ---- P = softmax(dim=2) * softmax(dim=1), in place (MU:230) ------------------------------
    for (int e = tid; e < nr * Cn; e += 1024) {
        const int i = e / Cn, j = e - i * Cn;
        const float v = S[e];
        S[e] = __fdiv_rn(expf(v - rmax[i]), rinv[i]) * __fdiv_rn(expf(v - cmax[j]), cinv[j]);
    }
    __syncthreads();
    // ---- foreground masks (MU:231-234) --------------------------------------------------------
    for (int i = warp_id(); i < nr; i += 32) {
        float m = -INFINITY;
        for (int j = 1 + lane_id(); j < Cn; j += 32) m = fmaxf(m, S[i * Cn + j]);
        m = warp_max(m);
        if (lane_id() == 0) {
            const float f = (m > S[i * Cn]) ? 1.f : 0.f;
            w1s[i] = f;
            if (r0 + i >= 1) w1[(size_t)b * N1 + r0 + i - 1] = f;
        }
    }
    for (int j = warp_id(); j < Cn; j += 32) {  // column max over rows >= 1 of this slab; row 0 itself
        float m = -INFINITY;
        for (int i = lane_id(); i < nr; i += 32)
            if (r0 + i >= 1) m = fmaxf(m, S[i * Cn + j]);
        m = warp_max(m);
        if (lane_id() < kAsgCl) {
            *cluster.map_shared_rank(&xch[rank * Cn + j], lane_id()) = m;
            if (rank == 0) *cluster.map_shared_rank(&p0[j], lane_id()) = S[j];
        }
    }
    cluster.sync();
    for (int j = tid; j < Cn; j += 1024) {
        float m = xch[j];
#pragma unroll
        for (int q = 1; q < kAsgCl; ++q) m = fmaxf(m, xch[q * Cn + j]);
        const float f = (j >= 1 && m > p0[j]) ? 1.f : 0.f;
        w2s[j] = f;
        if (rank == 0 && j >= 1) w2[(size_t)b * N2 + j - 1] = f;
    }
    __syncthreads();
    // ---- score**1.5 -> fp64 prefix sum rounded to fp32 per element -> / (last + 1e-8) ----------
    // this slab's share of the flattened [1:,1:] block: rows max(r0,1) .. r0+nr-1
    const int i_lo = max(r0, 1) - r0;                 // first local row that belongs to the block
    const int n_el = max(0, nr - i_lo) * N2;          // elements of this slab
    const int chunk = (n_el + 1023) / 1024;
    const int e0 = min(n_el, tid * chunk), e1 = min(n_el, e0 + chunk);
    auto q_at = [&](int il, int j) { return pow15(S[il * Cn + j + 1] * w1s[il] * w2s[j + 1]); };
    double local = 0.0;
    {
        int il = i_lo + e0 / N2, j = e0 - (e0 / N2) * N2;
        for (int e = e0; e < e1; ++e) {
            local += (double)q_at(il, j);
            if (++j == N2) { j = 0; ++il; }
        }
    }
    double inc = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double nb = __shfl_up_sync(kFull, inc, o);
        if (lane_id() >= o) inc += nb;
    }
    if (lane_id() == 31) s_scan[warp_id()] = inc;
    __syncthreads();
    if (warp_id() == 0) {
        const double v = s_scan[lane_id()];
        double winc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double nb = __shfl_up_sync(kFull, winc, o);
            if (lane_id() >= o) winc += nb;
        }
        s_scan[lane_id()] = winc - v;
        if (lane_id() == 31) s_scan[32] = winc;  // slab total
    }
    __syncthreads();
    if (tid < kAsgCl) *cluster.map_shared_rank(&s_tot[rank], tid) = s_scan[32];
    cluster.sync();
    double offset = 0.0, total = 0.0;
#pragma unroll
    for (int q = 0; q < kAsgCl; ++q) {
        if (q < rank) offset += s_tot[q];
        total += s_tot[q];
    }
    double run = offset + s_scan[warp_id()] + (inc - local);
    const float denom = (float)total + 1e-8f;
    float *out = cdf + (size_t)b * N1 * N2 + (size_t)(max(r0, 1) - 1) * N2;
    {
        int il = i_lo + e0 / N2, j = e0 - (e0 / N2) * N2;
        for (int e = e0; e < e1; ++e) {
            run += (double)q_at(il, j);
            out[e] = __fdiv_rn((float)run, denom);
            if (++j == N2) { j = 0; ++il; }
        }
    }
    cluster.sync();  // keep shared memory alive until every peer's remote stores have landed
}

static size_t coarse_assign_cluster_smem(int N1, int N2) {
    const size_t R = N1 + 1, Cn = N2 + 1, per = (R + kAsgCl - 1) / kAsgCl;
    return (34 + kAsgCl) * sizeof(double) + (per * Cn + 2 * kAsgCl * Cn + 3 * per + 4 * Cn) * sizeof(float);
}

static size_t coarse_assign_smem(int N1, int N2) {
    const size_t R = N1 + 1, Cn = N2 + 1;
    return 34 * sizeof(double) + (R * Cn + 3 * R + 3 * Cn) * sizeof(float);
}

// ------------------------------------------------------------------------------------
// coarse_hypotheses: one thread per hypothesis
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
coarse_hypotheses_kernel(const float *__restrict__ cdf, const float *__restrict__ rnd,
                         const float *__restrict__ pts1, const float *__restrict__ pts2, int N1,
                         int N2, int P1, float *__restrict__ Rout, float *__restrict__ tout,
                         float *__restrict__ resid, int32_t *__restrict__ corr) {
    // two-level lower bound: a shared-memory index holds every kStride-th cdf entry, so only the
    // last log2(kStride) probes (one or two cache lines) go to global memory
    constexpr int kStride = 64, kMaxIdx = 1024;
    __shared__ float s_idx[kMaxIdx];
    const int b = blockIdx.y, h = blockIdx.x * blockDim.x + threadIdx.x;
    const int E = N1 * N2;
    const float *c = cdf + (size_t)b * E;
    const int n_idx = min(kMaxIdx, (E + kStride - 1) / kStride);  // entry k = c[min(E, (k+1)*kStride) - 1]
    for (int k = threadIdx.x; k < n_idx; k += blockDim.x) s_idx[k] = __ldg(c + min(E, (k + 1) * kStride) - 1);
    __syncthreads();
    if (h >= P1) return;
    const size_t hb = (size_t)b * P1 + h;
    float p1[3][3], p2[3][3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float r = __ldg(rnd + hb * 3 + d);
        // first block whose last element is >= r (blocks beyond the index are searched directly)
        int blo = 0, bhi = n_idx;
        while (blo < bhi) {
            const int mid = (blo + bhi) >> 1;
            if (s_idx[mid] >= r) bhi = mid; else blo = mid + 1;
        }
        int lo = blo * kStride, hi = (blo < n_idx) ? min(E, (blo + 1) * kStride) - 1 : E;
        // invariant: c[hi] >= r when blo < n_idx (hi is a valid answer), everything before lo is < r
        while (lo < hi) {  // first e with c[e] >= r (MU:277-305); none -> 0
            const int mid = (lo + hi) >> 1;
            if (__ldg(c + mid) >= r) hi = mid; else lo = mid + 1;
        }
        const int e = lo < E ? lo : 0;
        if (corr) corr[hb * 3 + d] = e;
        const int i1 = min(e / N2, N1 - 1), i2 = min(e % N2, N2 - 1);
        const float *a = pts1 + ((size_t)b * N1 + i1) * 3, *q = pts2 + ((size_t)b * N2 + i2) * 3;
#pragma unroll
        for (int k = 0; k < 3; ++k) { p1[d][k] = __ldg(a + k); p2[d][k] = __ldg(q + k); }
    }
    // weighted_procrustes(src = p2, ref = p1, weights = None) with weight_thresh 0.5:
    // every weight is 1 / (3 + 1e-5) (MU:380-384)
    const float w = __fdiv_rn(1.0f, 3.0f + 1e-5f);
    float cs[3], cr[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        cs[k] = (p2[0][k] * w + p2[1][k] * w) + p2[2][k] * w;
        cr[k] = (p1[0][k] * w + p1[1][k] * w) + p1[2][k] * w;
    }
    float H[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < 3; ++d) s = fmaf(p2[d][a] - cs[a], w * (p1[d][k] - cr[k]), s);
            H[a][k] = s;
        }
    float Rm[3][3];
    kabsch_rotation(H, Rm);
    float t[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
        t[a] = cr[a] - (Rm[a][0] * cs[0] + Rm[a][1] * cs[1] + Rm[a][2] * cs[2]);
    // residual: mean_d || (p1_d - t) R - p2_d ||  (MU:261)
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float v0 = p1[d][0] - t[0], v1 = p1[d][1] - t[1], v2 = p1[d][2] - t[2];
        float n2 = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float o = v0 * Rm[0][k] + v1 * Rm[1][k] + v2 * Rm[2][k] - p2[d][k];
            n2 = fmaf(o, o, n2);
        }
        acc += sqrtf(n2);
    }
    resid[hb] = acc / 3.0f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int k = 0; k < 3; ++k) Rout[hb * 9 + a * 3 + k] = Rm[a][k];
        tout[hb * 3 + a] = t[a];
    }
}

// ------------------------------------------------------------------------------------
// select_smallest: radix select (4 x 8 bits) + ordered compaction, one CTA per instance
// ------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ordered_key(float v) {
    unsigned u = __float_as_uint(v);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;  // NaN sorts last
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(1024, 1)
select_smallest_kernel(const float *__restrict__ resid, int P1, int P2, int32_t *__restrict__ sel) {
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_k;
    __shared__ int s_cnt[2][33];
    const int tid = threadIdx.x, b = blockIdx.x;
    const float *v = resid + (size_t)b * P1;
    if (tid == 0) { s_prefix = 0u; s_k = (unsigned)P2; }
    unsigned mask = 0u;
    for (int pass = 3; pass >= 0; --pass) {
        if (tid < 256) hist[tid] = 0u;
        __syncthreads();
        const unsigned prefix = s_prefix;
        for (int e = tid; e < P1; e += 1024) {
            const unsigned u = ordered_key(__ldg(v + e));
            if ((u & mask) == prefix) atomicAdd(&hist[(u >> (8 * pass)) & 255u], 1u);
        }
        __syncthreads();
        if (warp_id() == 0) {  // which of the 256 bins holds the k-th element: 8 bins per lane
            const unsigned k = s_k;
            unsigned loc = 0u;
#pragma unroll
            for (int q = 0; q < 8; ++q) loc += hist[lane_id() * 8 + q];
            unsigned inc = loc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned nb = __shfl_up_sync(kFull, inc, o);
                if (lane_id() >= o) inc += nb;
            }
            const unsigned before = inc - loc;  // elements in lower lanes' bins
            const unsigned hit = __ballot_sync(kFull, inc >= k);  // first lane whose cumulative count reaches k
            const int owner = hit ? __ffs(hit) - 1 : 31;
            if (lane_id() == owner) {
                unsigned cum = before, bin = lane_id() * 8;
                for (int q = 0; q < 7; ++q, ++bin) {
                    if (cum + hist[bin] >= k) break;
                    cum += hist[bin];
                }
                s_k = k - cum;
                s_prefix = prefix | (bin << (8 * pass));
            }
        }
        mask |= 255u << (8 * pass);
        __syncthreads();
    }
    const unsigned T = s_prefix;     // key of the P2-th smallest element
    const int need = (int)s_k;       // how many elements equal to T belong to the set
    const int chunk = (P1 + 1023) / 1024;
    const int e0 = min(P1, tid * chunk), e1 = min(P1, e0 + chunk);
    int nl = 0, ne = 0;
    for (int e = e0; e < e1; ++e) {
        const unsigned u = ordered_key(__ldg(v + e));
        nl += (u < T);
        ne += (u == T);
    }
    // block exclusive scan of (nl, ne)
    int il = nl, ie = ne;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(kFull, il, o), c = __shfl_up_sync(kFull, ie, o);
        if (lane_id() >= o) { il += a; ie += c; }
    }
    if (lane_id() == 31) { s_cnt[0][warp_id()] = il; s_cnt[1][warp_id()] = ie; }
    __syncthreads();
    if (warp_id() == 0) {
        const int a = s_cnt[0][lane_id()], c = s_cnt[1][lane_id()];
        int wa = a, wc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(kFull, wa, o), y = __shfl_up_sync(kFull, wc, o);
            if (lane_id() >= o) { wa += x; wc += y; }
        }
        s_cnt[0][lane_id()] = wa - a;
        s_cnt[1][lane_id()] = wc - c;
    }
    __syncthreads();
    int lb = s_cnt[0][warp_id()] + il - nl, eb = s_cnt[1][warp_id()] + ie - ne;
    int32_t *out = sel + (size_t)b * P2;
    for (int e = e0; e < e1; ++e) {
        const unsigned u = ordered_key(__ldg(v + e));
        if (u < T) {
            out[lb + min(eb, need)] = e;
            ++lb;
        } else if (u == T) {
            if (eb < need) out[lb + eb] = e;
            ++eb;
        }
    }
}

// ------------------------------------------------------------------------------------
// score_hypotheses
// ------------------------------------------------------------------------------------
constexpr int kScHypsMax = 8;  // hypotheses per CTA (8 for large sweeps, 4 to fill the GPU on small ones)
constexpr int kScQMax = 8;     // register-held query points per thread

template <int Q, int kScThreads>
__device__ __forceinline__ void score_items(const float4 *__restrict__ s_model, int Nm,
                                            const float *__restrict__ s_Rt,
                                            const int *__restrict__ s_fg,
                                            const float *__restrict__ pts1_b, int nfg, int items,
                                            float *__restrict__ s_dis) {
    float qx[Q], qy[Q], qz[Q], q2[Q], best[Q];
#pragma unroll
    for (int u = 0; u < Q; ++u) {
        const int it = threadIdx.x + u * kScThreads;
        qx[u] = qy[u] = qz[u] = q2[u] = 0.f;
        best[u] = INFINITY;
        if (it < items) {
            const int hh = it / nfg, qi = s_fg[it - hh * nfg];
            const float *Rt = s_Rt + hh * 12;
            const float *p = pts1_b + 3 * qi;
            const float v0 = __ldg(p) - Rt[9], v1 = __ldg(p + 1) - Rt[10], v2 = __ldg(p + 2) - Rt[11];
            // (p1 - t) @ R : row vector times R (MU:267)
            qx[u] = v0 * Rt[0] + v1 * Rt[3] + v2 * Rt[6];
            qy[u] = v0 * Rt[1] + v1 * Rt[4] + v2 * Rt[7];
            qz[u] = v0 * Rt[2] + v1 * Rt[5] + v2 * Rt[8];
            q2[u] = qx[u] * qx[u] + qy[u] * qy[u] + qz[u] * qz[u];
        }
    }
    // two model points per trip: 6 FFMA + ONE 3-input minimum (FMNMX3) per query
    int m = 0;
#pragma unroll 2
    for (; m + 1 < Nm; m += 2) {
        const float4 a = s_model[m], c = s_model[m + 1];  // (-2mx, -2my, -2mz, |m|^2), broadcast
#pragma unroll
        for (int u = 0; u < Q; ++u) {
            const float d0 = fmaf(qx[u], a.x, fmaf(qy[u], a.y, fmaf(qz[u], a.z, a.w)));
            const float d1 = fmaf(qx[u], c.x, fmaf(qy[u], c.y, fmaf(qz[u], c.z, c.w)));
            best[u] = fminf(fminf(best[u], d0), d1);
        }
    }
    if (m < Nm) {
        const float4 a = s_model[m];
#pragma unroll
        for (int u = 0; u < Q; ++u)
            best[u] = fminf(best[u], fmaf(qx[u], a.x, fmaf(qy[u], a.y, fmaf(qz[u], a.z, a.w))));
    }
#pragma unroll
    for (int u = 0; u < Q; ++u) {
        const int it = threadIdx.x + u * kScThreads;
        if (it < items) s_dis[it] = sqrtf(fmaxf(q2[u] + best[u], 0.f));  // clamp, then sqrt
    }
}

template <int kScHyps, int kScThreads>
__global__ void __launch_bounds__(kScThreads)
score_hypotheses_kernel(const float *__restrict__ R, const float *__restrict__ t,
                        const int32_t *__restrict__ sel, int P, int H,
                        const float *__restrict__ pts1, const float *__restrict__ w1, int N1,
                        const float *__restrict__ model, int Nm, float *__restrict__ scores) {
    extern __shared__ float4 s_dyn4[];
    float4 *s_model = s_dyn4;                                    // Nm
    float *s_dis = reinterpret_cast<float *>(s_model + Nm);      // kScHyps * N1
    float *s_Rt = s_dis + kScHyps * N1;                          // kScHyps * 12
    int *s_fg = reinterpret_cast<int *>(s_Rt + kScHyps * 12);    // N1
    __shared__ int s_nfg;
    const int tid = threadIdx.x, b = blockIdx.y, h0 = blockIdx.x * kScHyps;
    const int nh = min(kScHyps, H - h0);

    const float *mb = model + (size_t)b * Nm * 3;
    for (int m = tid; m < Nm; m += kScThreads) {
        const float x = __ldg(mb + 3 * m), y = __ldg(mb + 3 * m + 1), z = __ldg(mb + 3 * m + 2);
        s_model[m] = make_float4(-2.f * x, -2.f * y, -2.f * z, x * x + y * y + z * z);
    }
    if (tid < nh * 12) {
        const int hh = tid / 12, k = tid - hh * 12;
        const int src = sel ? __ldg(sel + (size_t)b * H + h0 + hh) : (h0 + hh);
        const size_t hb = (size_t)b * P + src;
        s_Rt[tid] = k < 9 ? __ldg(R + hb * 9 + k) : __ldg(t + hb * 3 + (k - 9));
    }
    if (warp_id() == 0) {  // ordered compaction of the foreground points
        const float *wb = w1 + (size_t)b * N1;
        int n = 0;
        for (int base = 0; base < N1; base += 32) {
            const int i = base + lane_id();
            const bool fg = i < N1 && __ldg(wb + i) != 0.f;
            const unsigned bal = __ballot_sync(kFull, fg);
            if (fg) s_fg[n + __popc(bal & ((1u << lane_id()) - 1u))] = i;
            n += __popc(bal);
        }
        if (lane_id() == 0) s_nfg = n;
    }
    __syncthreads();
    const int nfg = s_nfg, items = nh * nfg;
    const float *pb = pts1 + (size_t)b * N1 * 3;
    const int q = (items + kScThreads - 1) / kScThreads;  // block-uniform, <= kScQMax
    switch (q) {
        case 0: break;
        case 1: score_items<1, kScThreads>(s_model, Nm, s_Rt, s_fg, pb, nfg, items, s_dis); break;
        case 2: score_items<2, kScThreads>(s_model, Nm, s_Rt, s_fg, pb, nfg, items, s_dis); break;
        case 3: score_items<3, kScThreads>(s_model, Nm, s_Rt, s_fg, pb, nfg, items, s_dis); break;
        case 4: score_items<4, kScThreads>(s_model, Nm, s_Rt, s_fg, pb, nfg, items, s_dis); break;
        case 5: score_items<5, kScThreads>(s_model, Nm, s_Rt, s_fg, pb, nfg, items, s_dis); break;
        case 6: score_items<6, kScThreads>(s_model, Nm, s_Rt, s_fg, pb, nfg, items, s_dis); break;
        case 7: score_items<7, kScThreads>(s_model, Nm, s_Rt, s_fg, pb, nfg, items, s_dis); break;
        default: score_items<8, kScThreads>(s_model, Nm, s_Rt, s_fg, pb, nfg, items, s_dis); break;
    }
    __syncthreads();
    // deterministic per-hypothesis sum: warp hh sums its nfg distances in a fixed order
    for (int hh = warp_id(); hh < nh; hh += kScThreads / 32) {
        float s = 0.f;
        for (int i = lane_id(); i < nfg; i += 32) s += s_dis[hh * nfg + i];
        s = warp_sum(s);
        if (lane_id() == 0)
            scores[(size_t)b * H + h0 + hh] = __fdiv_rn((float)nfg, s + 1e-8f);  // MU:271
    }
}

static size_t score_smem(int N1, int Nm, int hyps) {
    return (size_t)Nm * sizeof(float4) + ((size_t)hyps * N1 + hyps * 12) * sizeof(float) +
           (size_t)N1 * sizeof(int);
}

// ------------------------------------------------------------------------------------
// select_best: first maximum of scores, one CTA per instance
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
select_best_kernel(const float *__restrict__ scores, const int32_t *__restrict__ sel,
                   const float *__restrict__ R, const float *__restrict__ t, int P, int H,
                   float *__restrict__ Rout, float *__restrict__ tout,
                   int32_t *__restrict__ proposal_id) {
    __shared__ unsigned long long s_key[8];
    __shared__ int s_win;
    const int b = blockIdx.x, tid = threadIdx.x;
    unsigned long long best = 0ull;
    for (int h = tid; h < H; h += 256) {
        const float v = __ldg(scores + (size_t)b * H + h);
        unsigned u = ordered_key(v);
        if (u == 0xffffffffu) u = 0u;  // NaN never wins
        const unsigned long long key = ((unsigned long long)u << 32) | (unsigned)(~h);
        best = key > best ? key : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long n = __shfl_xor_sync(kFull, best, o);
        best = n > best ? n : best;
    }
    if (lane_id() == 0) s_key[warp_id()] = best;
    __syncthreads();
    if (tid == 0) {
        unsigned long long k = s_key[0];
        for (int w = 1; w < 8; ++w) k = s_key[w] > k ? s_key[w] : k;
        s_win = (int)~(unsigned)(k & 0xffffffffull);
    }
    __syncthreads();
    const int win = s_win;
    const int src = sel ? sel[(size_t)b * H + win] : win;
    const size_t hb = (size_t)b * P + src;
    if (tid < 9) Rout[(size_t)b * 9 + tid] = R[hb * 9 + tid];
    else if (tid < 12) tout[(size_t)b * 3 + tid - 9] = t[hb * 3 + tid - 9];
    else if (tid == 12 && proposal_id) proposal_id[b] = src;
}

}  // namespace pem

// ------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------
using namespace pem;

extern "C" int pem_coarse_assign(const float *atten, int B, int N1, int N2, float *w1, float *w2,
                                 float *cdf, pem_stream_t stream) {
    PEM_REQUIRE_PTR(atten);
    PEM_REQUIRE_PTR(w1);
    PEM_REQUIRE_PTR(w2);
    PEM_REQUIRE_PTR(cdf);
    PEM_REQUIRE(B > 0 && N1 > 0 && N2 > 0, PEM_E_SHAPE);
    if (N1 + 1 >= 8 * kAsgCl && (long long)((N1 + kAsgCl) / kAsgCl) * (N2 + 1) <= 40000 &&
        coarse_assign_cluster_smem(N1, N2) <= (size_t)227 * 1024) {
        // thread-block cluster: kAsgCl SMs per instance (shapes whose per-CTA slice plus the 2*kAsgCl
        // exchange rows exceed 227 KB fall through to the single-CTA kernel or PEM_E_UNSUPPORTED)
        const size_t smem = coarse_assign_cluster_smem(N1, N2);
        static PemSmemCache cached(0);
        if (int rc = pem_ensure_dyn_smem(coarse_assign_cluster_kernel, smem, &cached)) return rc;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(B * kAsgCl);
        cfg.blockDim = dim3(1024);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = pem_cu_stream(stream);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kAsgCl;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, coarse_assign_cluster_kernel, atten, N1, N2, w1, w2, cdf);
        if (e != cudaSuccess) return (int)e;
        return pem_launch_status();
    }
    PEM_REQUIRE((long long)(N1 + 1) * (N2 + 1) <= 50176 && coarse_assign_smem(N1, N2) <= (size_t)227 * 1024,
                PEM_E_UNSUPPORTED);
    const size_t smem = coarse_assign_smem(N1, N2);
    static PemSmemCache cached(0);
    if (int rc = pem_ensure_dyn_smem(coarse_assign_kernel, smem, &cached)) return rc;
    coarse_assign_kernel<<<B, 1024, smem, pem_cu_stream(stream)>>>(atten, N1, N2, w1, w2, cdf);
    return pem_launch_status();
}

extern "C" int pem_coarse_hypotheses(const float *cdf, const float *rnd, const float *pts1,
                                     const float *pts2, int B, int N1, int N2, int P1, float *R,
                                     float *t, float *resid, int32_t *corr, pem_stream_t stream) {
    PEM_REQUIRE_PTR(cdf);
    PEM_REQUIRE_PTR(rnd);
    PEM_REQUIRE_PTR(pts1);
    PEM_REQUIRE_PTR(pts2);
    PEM_REQUIRE_PTR(R);
    PEM_REQUIRE_PTR(t);
    PEM_REQUIRE_PTR(resid);
    PEM_REQUIRE(B > 0 && N1 > 0 && N2 > 0 && P1 > 0 && B <= 65535, PEM_E_SHAPE);
    coarse_hypotheses_kernel<<<dim3((P1 + 127) / 128, B), 128, 0, pem_cu_stream(stream)>>>(
        cdf, rnd, pts1, pts2, N1, N2, P1, R, t, resid, corr);
    return pem_launch_status();
}

extern "C" int pem_select_smallest(const float *resid, int B, int P1, int P2, int32_t *sel,
                                   pem_stream_t stream) {
    PEM_REQUIRE_PTR(resid);
    PEM_REQUIRE_PTR(sel);
    PEM_REQUIRE(B > 0 && P1 > 0 && P2 > 0 && P2 <= P1, PEM_E_SHAPE);
    select_smallest_kernel<<<B, 1024, 0, pem_cu_stream(stream)>>>(resid, P1, P2, sel);
    return pem_launch_status();
}

namespace pem {
int score_tc_launch(const float *R, const float *t, const int32_t *sel, int B, int P, int H,
                    const float *pts1, const float *w1, int N1, const float *model, int Nm,
                    float *scores, cudaStream_t st);  // score_tc.cu
}
extern "C" int pem_score_hypotheses(const float *R, const float *t, const int32_t *sel, int B, int P,
                                    int H, const float *pts1, const float *w1, int N1,
                                    const float *model, int Nm, float *scores,
                                    pem_stream_t stream) {
    return pem_score_hypotheses_ex(R, t, sel, B, P, H, pts1, w1, N1, model, Nm, scores, PEM_SCORE_AUTO, stream);
}

extern "C" int pem_score_hypotheses_ex(const float *R, const float *t, const int32_t *sel, int B, int P,
                                       int H, const float *pts1, const float *w1, int N1,
                                       const float *model, int Nm, float *scores, int kernel,
                                       pem_stream_t stream) {
    PEM_REQUIRE_PTR(R);
    PEM_REQUIRE_PTR(t);
    PEM_REQUIRE_PTR(pts1);
    PEM_REQUIRE_PTR(w1);
    PEM_REQUIRE_PTR(model);
    PEM_REQUIRE_PTR(scores);
    PEM_REQUIRE(B > 0 && P > 0 && H > 0 && N1 > 0 && Nm > 0 && B <= 65535, PEM_E_SHAPE);
    PEM_REQUIRE(sel != nullptr || H <= P, PEM_E_SHAPE);
    PEM_REQUIRE(N1 <= 256, PEM_E_UNSUPPORTED);  // hyps*N1 <= kScQMax*threads in both shapes
    PEM_REQUIRE(kernel == PEM_SCORE_AUTO || kernel == PEM_SCORE_SIMT || kernel == PEM_SCORE_TENSOR, PEM_E_SHAPE);
    if (kernel != PEM_SCORE_SIMT) {  // tensor-core kernel (score_tc.cu) for every shape it takes
        const int rc = score_tc_launch(R, t, sel, B, P, H, pts1, w1, N1, model, Nm, scores, pem_cu_stream(stream));
        if (rc != PEM_E_UNSUPPORTED || kernel == PEM_SCORE_TENSOR) return rc;
    }
    // 8 hypotheses per CTA amortise the model-point staging best; small problems use 4 so the
    // grid still covers every SM several times
    const bool big = (long long)B * ((H + 7) / 8) >= 4LL * PEM_NUM_SMS * 4;
    const size_t smem = score_smem(N1, Nm, big ? 8 : 4);
    PEM_REQUIRE(smem <= 227 * 1024, PEM_E_UNSUPPORTED);
    if (big) {
        static PemSmemCache cached(0);
        if (int rc = pem_ensure_dyn_smem(score_hypotheses_kernel<8, 256>, smem, &cached)) return rc;
        score_hypotheses_kernel<8, 256><<<dim3((H + 7) / 8, B), 256, smem, pem_cu_stream(stream)>>>(
            R, t, sel, P, H, pts1, w1, N1, model, Nm, scores);
    } else {
        static PemSmemCache cached(0);
        // 4 hypotheses x <=196 points over 128 threads: up to 7 register-held queries per thread
        if (int rc = pem_ensure_dyn_smem(score_hypotheses_kernel<4, 128>, smem, &cached)) return rc;
        score_hypotheses_kernel<4, 128><<<dim3((H + 3) / 4, B), 128, smem, pem_cu_stream(stream)>>>(
            R, t, sel, P, H, pts1, w1, N1, model, Nm, scores);
    }
    return pem_launch_status();
}

extern "C" int pem_select_best(const float *scores, const int32_t *sel, const float *R,
                               const float *t, int B, int P, int H, float *R_out, float *t_out,
                               int32_t *proposal_id, pem_stream_t stream) {
    PEM_REQUIRE_PTR(scores);
    PEM_REQUIRE_PTR(R);
    PEM_REQUIRE_PTR(t);
    PEM_REQUIRE_PTR(R_out);
    PEM_REQUIRE_PTR(t_out);
    PEM_REQUIRE(B > 0 && P > 0 && H > 0, PEM_E_SHAPE);
    select_best_kernel<<<B, 256, 0, pem_cu_stream(stream)>>>(scores, sel, R, t, P, H, R_out, t_out,
                                                             proposal_id);
    return pem_launch_status();
}
