// The operator seam of the reference's pose solvers as standalone batched entry points:
//   CustomSVD / CustomSVDu / CustomSVDv  (utils/model_utils.py:469-510; forward = torch.svd;
//                                         plugin twin ov_pointnet2_op/custom_svd.cpp:103-180, Eigen BDCSVD)
//   CustomDet                            (utils/model_utils.py:512-526; forward = torch.det;
//                                         plugin twin ov_pointnet2_op/custom_det.cpp:52-74)
//   CustomSearchSorted                   (utils/model_utils.py:457-467; forward = torch.searchsorted;
//                                         plugin twin ov_pointnet2_op/custom_searchsorted.cpp:90-109)
// at the shapes the reference's own op scripts use (ov_op_test/ov_test_custom_svd.py:31-33: 42000 x 3 x 3).
// One thread per matrix / per draw, everything in registers; the fused kernels of the hot path
// (coarse_hypotheses_kernel, weighted_procrustes_kernel) carry the same arithmetic inline (svd3.cuh).
#include "pem_common.cuh"
#include "svd3.cuh"

namespace pem {

// Full SVD H = U diag(S) V^T of a row-major 3x3 fp32 matrix: one-sided (Hestenes) Jacobi on the columns
// in fp64, S descending, U and V orthonormal also for rank-deficient H (missing left vectors completed;
// H = 0 gives U = V = I like LAPACK).
__device__ __forceinline__ void svd3_full(const float *__restrict__ h, float *__restrict__ Uo, float *__restrict__ So,
                                          float *__restrict__ Vo) {
    double G[3][3], V[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            G[r][c] = (double)h[r * 3 + c];
            V[r][c] = (r == c) ? 1.0 : 0.0;
        }
#pragma unroll 1
    for (int sweep = 0; sweep < 6; ++sweep) {
        jacobi_pair(G, V, 0, 1);
        jacobi_pair(G, V, 0, 2);
        jacobi_pair(G, V, 1, 2);
    }
    double n[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) n[c] = G[0][c] * G[0][c] + G[1][c] * G[1][c] + G[2][c] * G[2][c];
#define PEM_SWAPCOLS(A, B)                                        \
    if (n[A] < n[B]) {                                            \
        double tmp = n[A]; n[A] = n[B]; n[B] = tmp;               \
        _Pragma("unroll") for (int r = 0; r < 3; ++r) {           \
            tmp = G[r][A]; G[r][A] = G[r][B]; G[r][B] = tmp;      \
            tmp = V[r][A]; V[r][A] = V[r][B]; V[r][B] = tmp;      \
        }                                                         \
    }
    PEM_SWAPCOLS(0, 1)
    PEM_SWAPCOLS(0, 2)
    PEM_SWAPCOLS(1, 2)
#undef PEM_SWAPCOLS
    double u[3][3];  // u[k] = k-th left singular vector
    const double tiny = 1e-28 * n[0];
    if (!(n[0] > 0.0)) {  // H == 0 (or NaN)
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                u[k][r] = (k == r) ? 1.0 : 0.0;
                V[r][k] = (k == r) ? 1.0 : 0.0;
            }
        n[0] = n[1] = n[2] = isnan(n[0]) ? n[0] : 0.0;
    } else {
        const double inv1 = rsqrt(n[0]);
#pragma unroll
        for (int r = 0; r < 3; ++r) u[0][r] = G[r][0] * inv1;
        // second vector: the column, re-orthogonalised against u1; rank 1 -> any unit vector orthogonal to u1
        double g2[3] = {G[0][1], G[1][1], G[2][1]};
        const double pr = g2[0] * u[0][0] + g2[1] * u[0][1] + g2[2] * u[0][2];
#pragma unroll
        for (int r = 0; r < 3; ++r) g2[r] -= pr * u[0][r];
        const double l2 = g2[0] * g2[0] + g2[1] * g2[1] + g2[2] * g2[2];
        if (l2 > tiny) {
            const double inv2 = rsqrt(l2);
#pragma unroll
            for (int r = 0; r < 3; ++r) u[1][r] = g2[r] * inv2;
        } else {
            const double a0 = fabs(u[0][0]), a1 = fabs(u[0][1]), a2 = fabs(u[0][2]);
            double e[3];
            e[0] = (a0 <= a1 && a0 <= a2) ? 1.0 : 0.0;
            e[1] = (e[0] == 0.0 && a1 <= a2) ? 1.0 : 0.0;
            e[2] = (e[0] == 0.0 && e[1] == 0.0) ? 1.0 : 0.0;
            cross3(u[0], e, u[1]);
            const double inv2 = rsqrt(u[1][0] * u[1][0] + u[1][1] * u[1][1] + u[1][2] * u[1][2]);
#pragma unroll
            for (int r = 0; r < 3; ++r) u[1][r] *= inv2;
        }
        // third vector: +-(u1 x u2), the sign that matches the third column when it carries one
        cross3(u[0], u[1], u[2]);
        const double dir = G[0][2] * u[2][0] + G[1][2] * u[2][1] + G[2][2] * u[2][2];
        if (n[2] > tiny && dir < 0.0) {
#pragma unroll
            for (int r = 0; r < 3; ++r) u[2][r] = -u[2][r];
        }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (Uo) Uo[r * 3 + c] = (float)u[c][r];
            if (Vo) Vo[r * 3 + c] = (float)V[r][c];
        }
    if (So) {
#pragma unroll
        for (int c = 0; c < 3; ++c) So[c] = (float)sqrt(n[c]);
    }
}

// 4 threads-worth of coalescing is not worth a transpose here: 36 B in, up to 84 B out per matrix, L2-resident
__global__ void __launch_bounds__(128)
svd3_batched_kernel(const float *__restrict__ H, long long n, float *__restrict__ U, float *__restrict__ S,
                    float *__restrict__ V) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float h[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) h[k] = __ldg(H + i * 9 + k);
    float u[9], s[3], v[9];
    svd3_full(h, u, s, v);
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        if (U) U[i * 9 + k] = u[k];
        if (V) V[i * 9 + k] = v[k];
    }
    if (S) {
#pragma unroll
        for (int k = 0; k < 3; ++k) S[i * 3 + k] = s[k];
    }
}

__global__ void __launch_bounds__(256)
det3_batched_kernel(const float *__restrict__ H, long long n, float *__restrict__ det) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) a[k] = (double)__ldg(H + i * 9 + k);
    // cofactor expansion in fp64, rounded once: the correctly rounded determinant of the fp32 entries up
    // to fp64 cancellation (LAPACK's fp32 LU lands within a few fp32 ulps of |a|^3 of it)
    const double d = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) +
                     a[2] * (a[3] * a[7] - a[4] * a[6]);
    det[i] = (float)d;
}

// out[b,s] = first e in [0,N] with seq[b,e] >= val[b,s] (N when there is none): torch.searchsorted(right=False)
// and custom_searchsorted.cpp:90-109.  Warp-cooperative first level over a strided shared-memory index
// is not needed at these sizes: the cdf row (N <= 64 K floats typically 38416) is L1/L2 resident.
__global__ void __launch_bounds__(256)
searchsorted_kernel(const float *__restrict__ seq, const float *__restrict__ val, int N, int S,
                    long long *__restrict__ out) {
    const int b = blockIdx.y, s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const float *c = seq + (size_t)b * N;
    const float r = __ldg(val + (size_t)b * S + s);
    int lo = 0, hi = N;
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (__ldg(c + mid) >= r) hi = mid; else lo = mid + 1;
    }
    out[(size_t)b * S + s] = lo;
}

}  // namespace pem

using namespace pem;

extern "C" int pem_svd3_batched(const float *H, long long n, float *U, float *S, float *V, pem_stream_t stream) {
    PEM_REQUIRE_PTR(H);
    PEM_REQUIRE(U != nullptr || S != nullptr || V != nullptr, PEM_E_NULL);
    PEM_REQUIRE(n > 0 && n <= (1ll << 31) * 128, PEM_E_SHAPE);
    svd3_batched_kernel<<<(unsigned)((n + 127) / 128), 128, 0, pem_cu_stream(stream)>>>(H, n, U, S, V);
    return pem_launch_status();
}

extern "C" int pem_det3_batched(const float *H, long long n, float *det, pem_stream_t stream) {
    PEM_REQUIRE_PTR(H);
    PEM_REQUIRE_PTR(det);
    PEM_REQUIRE(n > 0 && n <= (1ll << 31) * 256, PEM_E_SHAPE);
    det3_batched_kernel<<<(unsigned)((n + 255) / 256), 256, 0, pem_cu_stream(stream)>>>(H, n, det);
    return pem_launch_status();
}

extern "C" int pem_searchsorted(const float *sorted_seq, const float *values, int B, int N, int S, int64_t *out,
                                pem_stream_t stream) {
    PEM_REQUIRE_PTR(sorted_seq);
    PEM_REQUIRE_PTR(values);
    PEM_REQUIRE_PTR(out);
    PEM_REQUIRE(B > 0 && B <= 65535 && N > 0 && S > 0, PEM_E_SHAPE);
    searchsorted_kernel<<<dim3((S + 255) / 256, B), 256, 0, pem_cu_stream(stream)>>>(sorted_seq, values, N, S,
                                                                                     (long long *)out);
    return pem_launch_status();
}
