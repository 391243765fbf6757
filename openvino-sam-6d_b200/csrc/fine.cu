// Fine pose stage of the PEM hot path: compute_fine_Rt (utils/model_utils.py:308-341 of the
// reference, MU below) on a materialised attention matrix, plus the generic weighted
// Procrustes solve (MU:343-436).
//
// The (B, N1+1, N2+1) matrix (16.8 MB per instance at the default 2049 x 2049) is streamed
// four times, never rewritten and never copied:
//   pass A  row max, column max                          (atomicMax on order-preserving ints)
//   pass B  row sum-exp, per-slab column sum-exp partials (fixed-order finish: deterministic)
//   pass C  foreground labels: row/column maxima of P = softmax_row * softmax_col vs the
//           background row/column (MU:321-322)
//   pass D  per-row sums  sum_j P_ij w2_j  and  sum_j P_ij w2_j pts2_j   (MU:324-330)
// followed by a one-CTA-per-instance weighted Procrustes (fp64 accumulation, fp64 3x3 SVD)
// and a pose-score kernel (nearest model point per transformed observed point, MU:334-339).
// Every pass is HBM-bound: slabs of 32 rows x all columns, threads own columns so global
// reads are fully coalesced, row reductions go through warp shuffles + shared memory.
// The fused tcgen05 path that never materialises the matrix is in fine_tc.cu.
#include "pem_common.cuh"
#include "svd3.cuh"
#include "tc_ptx.cuh"  // packed fp32 pairs (FFMA2)

namespace pem {

constexpr int kFnThreads = 256;
constexpr int kFnRows = 32;  // rows per slab

__device__ __forceinline__ int float_to_ordered_int(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_int_to_float(int i) {
    return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff);
}

// block reduction over kFnRows per-thread row values at once: s_red[row][warp]
template <bool IS_MAX>
__device__ __forceinline__ void slab_row_reduce(float (&v)[kFnRows], float *s_red /* kFnRows*8 */,
                                                float *s_out /* kFnRows */) {
#pragma unroll
    for (int r = 0; r < kFnRows; ++r) {
        float x = IS_MAX ? warp_max(v[r]) : warp_sum(v[r]);
        if (lane_id() == 0) s_red[r * 8 + warp_id()] = x;
    }
    __syncthreads();
    if (threadIdx.x < kFnRows) {
        float x = s_red[threadIdx.x * 8];
#pragma unroll
        for (int w = 1; w < 8; ++w)
            x = IS_MAX ? fmaxf(x, s_red[threadIdx.x * 8 + w]) : x + s_red[threadIdx.x * 8 + w];
        s_out[threadIdx.x] = x;
    }
    __syncthreads();
}

// pass A ------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFnThreads)
fine_init_kernel(int *__restrict__ colmax_i, unsigned *__restrict__ colmaxP, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        colmax_i[i] = float_to_ordered_int(-INFINITY);
        colmaxP[i] = 0u;
    }
}

__global__ void __launch_bounds__(kFnThreads)
fine_max_kernel(const float *__restrict__ atten, int R, int Cn, float *__restrict__ rowmax,
                int *__restrict__ colmax_i) {
    __shared__ float s_red[kFnRows * 8], s_out[kFnRows];
    const int b = blockIdx.y, i0 = blockIdx.x * kFnRows, tid = threadIdx.x;
    const int nr = min(kFnRows, R - i0);
    const float *A = atten + ((size_t)b * R + i0) * Cn;
    float rm[kFnRows];
#pragma unroll
    for (int r = 0; r < kFnRows; ++r) rm[r] = -INFINITY;
    for (int j = tid; j < Cn; j += kFnThreads) {
        float cm = -INFINITY;
#pragma unroll
        for (int r = 0; r < kFnRows; ++r) {
            if (r < nr) {
                const float v = __ldg(A + (size_t)r * Cn + j);
                rm[r] = fmaxf(rm[r], v);
                cm = fmaxf(cm, v);
            }
        }
        atomicMax(colmax_i + (size_t)b * Cn + j, float_to_ordered_int(cm));
    }
    slab_row_reduce<true>(rm, s_red, s_out);
    if (tid < nr) rowmax[(size_t)b * R + i0 + tid] = s_out[tid];
}

// pass B ------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFnThreads)
fine_sum_kernel(const float *__restrict__ atten, int R, int Cn, const float *__restrict__ rowmax,
                const int *__restrict__ colmax_i, float *__restrict__ rowsum,
                float *__restrict__ colpart /* (B, nslab, Cn) */) {
    __shared__ float s_red[kFnRows * 8], s_out[kFnRows], s_rm[kFnRows];
    const int b = blockIdx.y, i0 = blockIdx.x * kFnRows, tid = threadIdx.x;
    const int nr = min(kFnRows, R - i0);
    if (tid < kFnRows) s_rm[tid] = tid < nr ? rowmax[(size_t)b * R + i0 + tid] : 0.f;
    __syncthreads();
    const float *A = atten + ((size_t)b * R + i0) * Cn;
    float rs[kFnRows];
#pragma unroll
    for (int r = 0; r < kFnRows; ++r) rs[r] = 0.f;
    for (int j = tid; j < Cn; j += kFnThreads) {
        const float cmx = ordered_int_to_float(colmax_i[(size_t)b * Cn + j]);
        float cs = 0.f;
#pragma unroll
        for (int r = 0; r < kFnRows; ++r) {
            if (r < nr) {
                const float v = __ldg(A + (size_t)r * Cn + j);
                rs[r] += __expf(v - s_rm[r]);
                cs += __expf(v - cmx);
            }
        }
        colpart[((size_t)b * gridDim.x + blockIdx.x) * Cn + j] = cs;
    }
    slab_row_reduce<false>(rs, s_red, s_out);
    if (tid < nr) rowsum[(size_t)b * R + i0 + tid] = s_out[tid];
}

__global__ void __launch_bounds__(kFnThreads)
fine_colsum_finish_kernel(const float *__restrict__ colpart, int nslab, int Cn,
                          float *__restrict__ colsum) {
    const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Cn) return;
    float s = 0.f;
    for (int k = 0; k < nslab; ++k) s += colpart[((size_t)b * nslab + k) * Cn + j];
    colsum[(size_t)b * Cn + j] = s;
}

// P_ij = softmax_row * softmax_col, the same two factors the reference multiplies (MU:320)
__device__ __forceinline__ float dual_p(float v, float rm, float rinv, float cm, float cinv) {
    return (__expf(v - rm) * rinv) * (__expf(v - cm) * cinv);
}

// pass C ------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFnThreads)
fine_label_kernel(const float *__restrict__ atten, int R, int Cn, const float *__restrict__ rowmax,
                  const float *__restrict__ rowsum, const int *__restrict__ colmax_i,
                  const float *__restrict__ colsum, float *__restrict__ w1 /* (B,R-1) */,
                  unsigned *__restrict__ colmaxP, float *__restrict__ P0 /* (B,Cn) row 0 of P */) {
    __shared__ float s_red[kFnRows * 8], s_out[kFnRows], s_rm[kFnRows], s_ri[kFnRows],
        s_pbg[kFnRows];
    const int b = blockIdx.y, i0 = blockIdx.x * kFnRows, tid = threadIdx.x;
    const int nr = min(kFnRows, R - i0);
    if (tid < kFnRows) {
        s_rm[tid] = tid < nr ? rowmax[(size_t)b * R + i0 + tid] : 0.f;
        s_ri[tid] = tid < nr ? 1.0f / rowsum[(size_t)b * R + i0 + tid] : 0.f;
    }
    __syncthreads();
    const float *A = atten + ((size_t)b * R + i0) * Cn;
    float rmx[kFnRows];
#pragma unroll
    for (int r = 0; r < kFnRows; ++r) rmx[r] = -1.f;
    for (int j = tid; j < Cn; j += kFnThreads) {
        const float cm = ordered_int_to_float(colmax_i[(size_t)b * Cn + j]);
        const float ci = 1.0f / colsum[(size_t)b * Cn + j];
        float cmx = -1.f;
#pragma unroll
        for (int r = 0; r < kFnRows; ++r) {
            if (r < nr) {
                const float p = dual_p(__ldg(A + (size_t)r * Cn + j), s_rm[r], s_ri[r], cm, ci);
                if (j == 0) s_pbg[r] = p;          // background column entry of this row
                else rmx[r] = fmaxf(rmx[r], p);    // max over columns 1..
                if (i0 + r == 0) P0[(size_t)b * Cn + j] = p;  // background row
                else cmx = fmaxf(cmx, p);          // max over rows 1..
            }
        }
        if (cmx >= 0.f) atomicMax(colmaxP + (size_t)b * Cn + j, __float_as_uint(cmx));
    }
    slab_row_reduce<true>(rmx, s_red, s_out);
    if (tid < nr && i0 + tid >= 1)
        w1[(size_t)b * (R - 1) + i0 + tid - 1] = s_out[tid] > s_pbg[tid] ? 1.f : 0.f;
}

// pass D ------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFnThreads)
fine_wsum_kernel(const float *__restrict__ atten, int R, int Cn, const float *__restrict__ rowmax,
                 const float *__restrict__ rowsum, const int *__restrict__ colmax_i,
                 const float *__restrict__ colsum, const unsigned *__restrict__ colmaxP,
                 const float *__restrict__ P0, const float *__restrict__ w1,
                 const float *__restrict__ pts2 /* (B,Cn-1,3) */, float *__restrict__ w2_out,
                 float *__restrict__ rowsumA /* (B,R-1) */, float *__restrict__ pred /* (B,R-1,3) */) {
    __shared__ float s_red[kFnRows * 8], s_out[4][kFnRows], s_rm[kFnRows], s_ri[kFnRows];
    const int b = blockIdx.y, i0 = blockIdx.x * kFnRows, tid = threadIdx.x;
    const int nr = min(kFnRows, R - i0);
    if (tid < kFnRows) {
        s_rm[tid] = tid < nr ? rowmax[(size_t)b * R + i0 + tid] : 0.f;
        s_ri[tid] = tid < nr ? 1.0f / rowsum[(size_t)b * R + i0 + tid] : 0.f;
    }
    __syncthreads();
    const float *A = atten + ((size_t)b * R + i0) * Cn;
    float a0[kFnRows], ax[kFnRows], ay[kFnRows], az[kFnRows];
#pragma unroll
    for (int r = 0; r < kFnRows; ++r) a0[r] = ax[r] = ay[r] = az[r] = 0.f;
    for (int j = 1 + tid; j < Cn; j += kFnThreads) {
        const size_t cj = (size_t)b * Cn + j;
        const bool fg = __uint_as_float(colmaxP[cj]) > P0[cj];  // label2 > 0 (MU:322)
        if (blockIdx.x == 0 && w2_out) w2_out[(size_t)b * (Cn - 1) + j - 1] = fg ? 1.f : 0.f;
        if (!fg) continue;
        const float cm = ordered_int_to_float(colmax_i[cj]);
        const float ci = 1.0f / colsum[cj];
        const float *q = pts2 + ((size_t)b * (Cn - 1) + j - 1) * 3;
        const float qx = __ldg(q), qy = __ldg(q + 1), qz = __ldg(q + 2);
#pragma unroll
        for (int r = 0; r < kFnRows; ++r) {
            if (r < nr) {
                const float p = dual_p(__ldg(A + (size_t)r * Cn + j), s_rm[r], s_ri[r], cm, ci);
                a0[r] += p;
                ax[r] = fmaf(p, qx, ax[r]);
                ay[r] = fmaf(p, qy, ay[r]);
                az[r] = fmaf(p, qz, az[r]);
            }
        }
    }
    slab_row_reduce<false>(a0, s_red, s_out[0]);
    slab_row_reduce<false>(ax, s_red, s_out[1]);
    slab_row_reduce<false>(ay, s_red, s_out[2]);
    slab_row_reduce<false>(az, s_red, s_out[3]);
    if (tid < nr && i0 + tid >= 1) {
        const size_t ri = (size_t)b * (R - 1) + i0 + tid - 1;
        const float f = w1[ri];  // label1 > 0
        const float s = s_out[0][tid] * f;
        rowsumA[ri] = s;  // assignment_score (MU:330)
        const float inv = 1.0f / (s + 1e-6f);  // MU:327
        pred[ri * 3 + 0] = s_out[1][tid] * f * inv;
        pred[ri * 3 + 1] = s_out[2][tid] * f * inv;
        pred[ri * 3 + 2] = s_out[3][tid] * f * inv;
    }
}

// weighted Procrustes: one CTA per problem -------------------------------------------------
__global__ void __launch_bounds__(256)
weighted_procrustes_kernel(const float *__restrict__ src, const float *__restrict__ ref,
                           const float *__restrict__ weights, int n, float thresh, float eps,
                           float *__restrict__ Rout, float *__restrict__ tout,
                           int *__restrict__ zero3 /* (B,3) scratch to clear, or null */) {
    __shared__ double s_red[32];
    __shared__ double s_vec[9][33];
    pdl_enter();
    const int b = blockIdx.x, tid = threadIdx.x;
    if (zero3 && tid < 3) zero3[3 * b + tid] = 0;
    const float *S = src + (size_t)b * n * 3, *Rf = ref + (size_t)b * n * 3;
    const float *W = weights ? weights + (size_t)b * n : nullptr;
    auto wt = [&](int i) {
        const float w = W ? __ldg(W + i) : 1.0f;
        return w < thresh ? 0.0f : w;  // MU:380
    };
    double sw = 0.0;
    for (int i = tid; i < n; i += 256) sw += (double)wt(i);
    sw = block_sum(sw, s_red);
    const float inv = 1.0f / ((float)sw + eps);  // weights / (sum + eps), MU:381
    double c[6] = {0, 0, 0, 0, 0, 0};
    for (int i = tid; i < n; i += 256) {
        const double w = (double)(wt(i) * inv);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            c[k] += w * (double)__ldg(S + 3 * i + k);
            c[3 + k] += w * (double)__ldg(Rf + 3 * i + k);
        }
    }
    block_sum_n<6>(c, s_vec);
    const float cs[3] = {(float)c[0], (float)c[1], (float)c[2]};
    const float cr[3] = {(float)c[3], (float)c[4], (float)c[5]};
    double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = tid; i < n; i += 256) {
        const float w = wt(i) * inv;
        float s3[3], r3[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            s3[k] = __ldg(S + 3 * i + k) - cs[k];
            r3[k] = w * (__ldg(Rf + 3 * i + k) - cr[k]);
        }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int k = 0; k < 3; ++k) h[a * 3 + k] += (double)s3[a] * (double)r3[k];
    }
    block_sum_n<9>(h, s_vec);
    if (tid == 0) {
        float H[3][3], Rm[3][3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int k = 0; k < 3; ++k) H[a][k] = (float)h[a * 3 + k];
        kabsch_rotation(H, Rm);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int k = 0; k < 3; ++k) Rout[(size_t)b * 9 + a * 3 + k] = Rm[a][k];
            tout[(size_t)b * 3 + a] =
                cr[a] - (Rm[a][0] * cs[0] + Rm[a][1] * cs[1] + Rm[a][2] * cs[2]);
        }
    }
}

// pose score (MU:334-339) ---------------------------------------------------------------
// CTA = 256 points x 4 quarters of the model, 512 threads: thread (lane point pair, quarter) scans
// its quarter for two points (one LDS.128 per 8 FP ops), the four partial minima meet in shared
// memory.  2048 points x 16 instances = 128 CTAs: one per SM, 4 warps per scheduler.
constexpr int kPsLanes = 128, kPsQ = 2, kPsPoints = kPsLanes * kPsQ, kPsParts = 4;
constexpr int kPsThreads = kPsLanes * kPsParts;

__global__ void __launch_bounds__(kPsThreads)
fine_pose_score_kernel(const float *__restrict__ pts1, const float *__restrict__ w1,
                       const float *__restrict__ Rm, const float *__restrict__ tv,
                       const float *__restrict__ model, int N1, int Nm, float dis_thres,
                       int *__restrict__ counters /* (B,3): hits, fg, ticket */,
                       float *__restrict__ score) {
    extern __shared__ float4 s_model[];
    __shared__ float s_best[kPsParts][kPsPoints];
    __shared__ int s_h[kPsLanes / 32], s_f[kPsLanes / 32];
    pdl_enter();
    const int b = blockIdx.y, tid = threadIdx.x;
    const int pt = tid & (kPsLanes - 1), part = tid / kPsLanes;  // part is warp-uniform
    const float *mb = model + (size_t)b * Nm * 3;
    for (int m = tid; m < Nm; m += kPsThreads) {
        const float x = __ldg(mb + 3 * m), y = __ldg(mb + 3 * m + 1), z = __ldg(mb + 3 * m + 2);
        s_model[m] = make_float4(-2.f * x, -2.f * y, -2.f * z, x * x + y * y + z * z);
    }
    float qx[kPsQ], qy[kPsQ], qz[kPsQ], q2[kPsQ], best[kPsQ];
    const float *R9 = Rm + (size_t)b * 9, *t3 = tv + (size_t)b * 3;
#pragma unroll
    for (int u = 0; u < kPsQ; ++u) {
        const int i = blockIdx.x * kPsPoints + u * kPsLanes + pt;
        qx[u] = qy[u] = qz[u] = q2[u] = 0.f;
        best[u] = INFINITY;
        if (i < N1) {
            const float *p = pts1 + ((size_t)b * N1 + i) * 3;
            const float v0 = __ldg(p) - t3[0], v1 = __ldg(p + 1) - t3[1], v2 = __ldg(p + 2) - t3[2];
            qx[u] = v0 * R9[0] + v1 * R9[3] + v2 * R9[6];
            qy[u] = v0 * R9[1] + v1 * R9[4] + v2 * R9[7];
            qz[u] = v0 * R9[2] + v1 * R9[5] + v2 * R9[8];
            q2[u] = qx[u] * qx[u] + qy[u] * qy[u] + qz[u] * qz[u];
        }
    }
    __syncthreads();
    const int per = (Nm + kPsParts - 1) / kPsParts, m0 = part * per, m1 = min(Nm, m0 + per);
    static_assert(kPsQ == 2, "the scan below works on one fp32 pair per thread");
    const uint64_t QX = tc::pack2(qx[0], qx[1]), QY = tc::pack2(qy[0], qy[1]), QZ = tc::pack2(qz[0], qz[1]);
#pragma unroll 4
    for (int m = m0; m < m1; ++m) {  // both queries per model point in three FFMA2 (same fp32 fmas as before)
        const float4 a = s_model[m];
        float d0, d1;
        tc::unpack2(tc::fma2(QX, tc::pack2(a.x, a.x),
                             tc::fma2(QY, tc::pack2(a.y, a.y), tc::fma2(QZ, tc::pack2(a.z, a.z), tc::pack2(a.w, a.w)))),
                    d0, d1);
        best[0] = fminf(best[0], d0);
        best[1] = fminf(best[1], d1);
    }
#pragma unroll
    for (int u = 0; u < kPsQ; ++u) s_best[part][u * kPsLanes + pt] = best[u];
    __syncthreads();
    if (part == 0) {  // warps 0..3
        int hits = 0, fgs = 0;
#pragma unroll
        for (int u = 0; u < kPsQ; ++u) {
            const int l = u * kPsLanes + pt, i = blockIdx.x * kPsPoints + l;
            int hit = 0, fg = 0;
            if (i < N1) {
                const float bm = fminf(fminf(s_best[0][l], s_best[1][l]), fminf(s_best[2][l], s_best[3][l]));
                fg = w1[(size_t)b * N1 + i] != 0.f;
                const float dis = sqrtf(fmaxf(q2[u] + bm, 0.f));
                hit = fg && (dis < dis_thres);
            }
            hits += __popc(__ballot_sync(kFull, hit));
            fgs += __popc(__ballot_sync(kFull, fg));
        }
        if (lane_id() == 0) { s_h[warp_id()] = hits; s_f[warp_id()] = fgs; }
    }
    __syncthreads();
    if (tid == 0) {
        int *c = counters + 3 * b;
        atomicAdd(c + 0, s_h[0] + s_h[1] + s_h[2] + s_h[3]);
        atomicAdd(c + 1, s_f[0] + s_f[1] + s_f[2] + s_f[3]);
        __threadfence();
        const int ticket = atomicAdd(c + 2, 1);
        if (ticket == (int)gridDim.x - 1) {  // last CTA of this instance finalises
            __threadfence();
            const float hits = (float)atomicAdd(c + 0, 0), nfg = (float)atomicAdd(c + 1, 0);
            // (hits / (sum(mask) + 1e-8)) * mean(mask)
            score[b] = __fdiv_rn(hits, nfg + 1e-8f) * __fdiv_rn(nfg, (float)N1);
        }
    }
}

struct FineWs {
    float *rowmax, *rowsum, *colsum, *P0, *colpart, *w1, *w2, *rowsumA, *pred;
    int *colmax_i;
    unsigned *colmaxP;
    int *counters;
    size_t bytes;
};

static FineWs fine_carve(void *base, int B, int N1, int N2) {
    const size_t R = (size_t)N1 + 1, Cn = (size_t)N2 + 1, nslab = (R + kFnRows - 1) / kFnRows;
    size_t off = 0;
    auto take = [&](size_t n_elems) {
        const size_t o = off;
        off += (n_elems * 4 + 255) / 256 * 256;
        return o;
    };
    char *p = (char *)base;
    FineWs w;
    const size_t o_rowmax = take(B * R), o_rowsum = take(B * R), o_colsum = take(B * Cn),
                 o_P0 = take(B * Cn), o_colpart = take(B * nslab * Cn), o_w1 = take(B * N1),
                 o_w2 = take(B * N2), o_rsA = take(B * N1), o_pred = take(B * N1 * 3),
                 o_cmi = take(B * Cn), o_cmP = take(B * Cn), o_cnt = take(B * 3);
    w.rowmax = (float *)(p + o_rowmax); w.rowsum = (float *)(p + o_rowsum);
    w.colsum = (float *)(p + o_colsum); w.P0 = (float *)(p + o_P0);
    w.colpart = (float *)(p + o_colpart); w.w1 = (float *)(p + o_w1); w.w2 = (float *)(p + o_w2);
    w.rowsumA = (float *)(p + o_rsA); w.pred = (float *)(p + o_pred);
    w.colmax_i = (int *)(p + o_cmi); w.colmaxP = (unsigned *)(p + o_cmP);
    w.counters = (int *)(p + o_cnt);
    w.bytes = off;
    return w;
}

// Shared tail of both fine paths: WSVD(pred_pts, pts1, assignment_score) with weight_thresh
// 0.0 (MU:319,331), then the pose score (MU:334-339).  counters: (B,3) ints of scratch.
int fine_tail_launch(const float *pred, const float *pts1, const float *rowsumA, const float *w1,
                     const float *model, int B, int N1, int Nm, float dis_thres, int *counters,
                     float *R, float *t, float *score, cudaStream_t st) {
    // (the Procrustes kernel also zeroes the pose-score counters of its instance)
    pem_launch_pdl(weighted_procrustes_kernel, dim3(B), dim3(256), 0, st, pred, pts1, rowsumA, N1, 0.0f, 1e-5f, R, t,
                   counters);
    const size_t smem = (size_t)Nm * sizeof(float4);
    static PemSmemCache cached(40 * 1024);
    if (int rc = pem_ensure_dyn_smem(fine_pose_score_kernel, smem, &cached)) return rc;
    pem_launch_pdl(fine_pose_score_kernel, dim3((N1 + kPsPoints - 1) / kPsPoints, B), dim3(kPsThreads), smem, st, pts1,
                   w1, R, t, model, N1, Nm, dis_thres, counters, score);
    return pem_launch_status();
}

}  // namespace pem

using namespace pem;

extern "C" size_t pem_fine_workspace_bytes(int B, int N1, int N2) {
    if (B <= 0 || N1 <= 0 || N2 <= 0) return 0;
    return fine_carve(nullptr, B, N1, N2).bytes;
}

extern "C" int pem_weighted_procrustes(const float *src, const float *ref, const float *weights,
                                       int Bp, int n, float weight_thresh, float eps, float *R,
                                       float *t, pem_stream_t stream) {
    PEM_REQUIRE_PTR(src);
    PEM_REQUIRE_PTR(ref);
    PEM_REQUIRE_PTR(R);
    PEM_REQUIRE_PTR(t);
    PEM_REQUIRE(Bp > 0 && n > 0, PEM_E_SHAPE);
    weighted_procrustes_kernel<<<Bp, 256, 0, pem_cu_stream(stream)>>>(src, ref, weights, n,
                                                                      weight_thresh, eps, R, t, nullptr);
    return pem_launch_status();
}

extern "C" int pem_fine_Rt(const float *atten, const float *pts1, const float *pts2,
                           const float *model, int B, int N1, int N2, int Nm, float dis_thres,
                           float *R, float *t, float *score, float *w1_out, float *w2_out,
                           float *rowsum_out, float *pred_out, void *workspace,
                           size_t workspace_bytes, pem_stream_t stream) {
    PEM_REQUIRE_PTR(atten);
    PEM_REQUIRE_PTR(pts1);
    PEM_REQUIRE_PTR(pts2);
    PEM_REQUIRE_PTR(model);
    PEM_REQUIRE_PTR(R);
    PEM_REQUIRE_PTR(t);
    PEM_REQUIRE_PTR(score);
    PEM_REQUIRE(B > 0 && N1 > 0 && N2 > 0 && Nm > 0 && B <= 65535, PEM_E_SHAPE);
    PEM_REQUIRE((size_t)Nm * sizeof(float4) <= 200 * 1024, PEM_E_UNSUPPORTED);
    PEM_REQUIRE(workspace != nullptr && workspace_bytes >= pem_fine_workspace_bytes(B, N1, N2),
                PEM_E_WORKSPACE);
    cudaStream_t st = pem_cu_stream(stream);
    FineWs w = fine_carve(workspace, B, N1, N2);
    float *w1 = w1_out ? w1_out : w.w1;
    float *rsA = rowsum_out ? rowsum_out : w.rowsumA;
    float *pred = pred_out ? pred_out : w.pred;
    const int R_ = N1 + 1, Cn = N2 + 1, nslab = (R_ + kFnRows - 1) / kFnRows;
    const dim3 grid(nslab, B);

    fine_init_kernel<<<(B * Cn + 255) / 256, 256, 0, st>>>(w.colmax_i, w.colmaxP, B * Cn);
    fine_max_kernel<<<grid, kFnThreads, 0, st>>>(atten, R_, Cn, w.rowmax, w.colmax_i);
    fine_sum_kernel<<<grid, kFnThreads, 0, st>>>(atten, R_, Cn, w.rowmax, w.colmax_i, w.rowsum,
                                                 w.colpart);
    fine_colsum_finish_kernel<<<dim3((Cn + 255) / 256, B), 256, 0, st>>>(w.colpart, nslab, Cn,
                                                                         w.colsum);
    fine_label_kernel<<<grid, kFnThreads, 0, st>>>(atten, R_, Cn, w.rowmax, w.rowsum, w.colmax_i,
                                                   w.colsum, w1, w.colmaxP, w.P0);
    fine_wsum_kernel<<<grid, kFnThreads, 0, st>>>(atten, R_, Cn, w.rowmax, w.rowsum, w.colmax_i,
                                                  w.colsum, w.colmaxP, w.P0, w1, pts2, w2_out, rsA,
                                                  pred);
    return fine_tail_launch(pred, pts1, rsA, w1, model, B, N1, Nm, dis_thres, w.counters, R, t, score,
                            st);
}
