// GeometricStructureEmbedding.get_embedding_indices (model/transformer.py:304-341 of the reference):
// the geometry that feeds the coarse transformer's structure embedding.
//
//   d_indices[b,i,j]   = sqrt(max(|p_i|^2 - 2 p_i.p_j + |p_j|^2, 0)) / sigma_d         (:321-322, MU:101-128)
//   knn(i)             = entries 1..k of the (k+1) smallest distances of row i            (:325)
//   a_indices[b,i,j,l] = atan2(max(|r x a|, 1e-8), clamp(r.a, -1, 1)) * 180/(sigma_a*pi)  (:333-343)
//                        with r = p_knn_l(i) - p_i and a = p_j - p_i
//
// The reference materialises (B,N,N,k,3) expansions of both vector sets (7 MB per instance at
// N = 197, ~20 launches); here one CTA per (instance, row i) keeps the cloud and the row of
// distances in shared memory, a warp picks the k+1 nearest by repeated (distance, index) argmin,
// and every thread finishes its columns j.  Ties between equal distances go to the lower index
// (torch.topk leaves them unspecified).
#include "pem_common.cuh"

namespace pem {

constexpr int kGeoThreads = 128;
constexpr int kGeoMaxK = 8;

__global__ void __launch_bounds__(kGeoThreads)
geo_embedding_kernel(const float *__restrict__ pts, int N, int k, float sigma_d, float factor_a,
                     float *__restrict__ d_idx /* (B,N,N) */, float *__restrict__ a_idx /* (B,N,N,k) */) {
    extern __shared__ float s_geo[];
    float *sx = s_geo, *sy = sx + N, *sz = sy + N, *sd = sz + N;  // cloud (SoA) and the distance row
    __shared__ int s_knn[kGeoMaxK + 1];
    const int b = blockIdx.y, i = blockIdx.x, tid = threadIdx.x;
    const float *pb = pts + (size_t)b * N * 3;
    for (int e = tid; e < 3 * N; e += kGeoThreads) {
        const int j = e / 3, c = e - 3 * j;
        s_geo[c * N + j] = __ldg(pb + e);
    }
    __syncthreads();
    const float xi = sx[i], yi = sy[i], zi = sz[i];
    const float x2 = __fadd_rn(__fadd_rn(__fmul_rn(xi, xi), __fmul_rn(yi, yi)), __fmul_rn(zi, zi));
    float *drow = d_idx + ((size_t)b * N + i) * N;
    for (int j = tid; j < N; j += kGeoThreads) {
        const float xj = sx[j], yj = sy[j], zj = sz[j];
        const float y2 = __fadd_rn(__fadd_rn(__fmul_rn(xj, xj), __fmul_rn(yj, yj)), __fmul_rn(zj, zj));
        const float xy = __fadd_rn(__fadd_rn(__fmul_rn(xi, xj), __fmul_rn(yi, yj)), __fmul_rn(zi, zj));
        const float sq = fmaxf(__fadd_rn(__fsub_rn(x2, __fmul_rn(2.f, xy)), y2), 0.f);  // x2 - 2xy + y2, clamped
        const float dist = sqrtf(sq);
        sd[j] = dist;
        drow[j] = __fdiv_rn(dist, sigma_d);
    }
    __syncthreads();
    if (warp_id() == 0) {  // the k+1 smallest of the row by (distance, index); the first one is dropped
        for (int r = 0; r <= k; ++r) {
            float best = INFINITY;
            int arg = 0x7fffffff;
            for (int j = lane_id(); j < N; j += 32) {
                const float v = sd[j];
                if (v < best) { best = v; arg = j; }  // ascending j: lowest index among equals
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(kFull, best, o);
                const int oa = __shfl_xor_sync(kFull, arg, o);
                if (ob < best || (ob == best && oa < arg)) { best = ob; arg = oa; }
            }
            if (lane_id() == 0) {
                s_knn[r] = arg < N ? arg : 0;
                if (arg < N) sd[arg] = INFINITY;  // taken
            }
            __syncwarp();
        }
    }
    __syncthreads();
    float rx[kGeoMaxK], ry[kGeoMaxK], rz[kGeoMaxK];
#pragma unroll
    for (int l = 0; l < kGeoMaxK; ++l) {
        if (l < k) {
            const int q = s_knn[l + 1];
            rx[l] = __fsub_rn(sx[q], xi); ry[l] = __fsub_rn(sy[q], yi); rz[l] = __fsub_rn(sz[q], zi);
        }
    }
    float *arow = a_idx + ((size_t)b * N + i) * N * k;
    for (int j = tid; j < N; j += kGeoThreads) {
        const float ax = __fsub_rn(sx[j], xi), ay = __fsub_rn(sy[j], yi), az = __fsub_rn(sz[j], zi);
#pragma unroll
        for (int l = 0; l < kGeoMaxK; ++l) {
            if (l < k) {
                // torch.cross / norm / sum in fp32, products and sums separately rounded
                const float cx = __fsub_rn(__fmul_rn(ry[l], az), __fmul_rn(rz[l], ay));
                const float cy = __fsub_rn(__fmul_rn(rz[l], ax), __fmul_rn(rx[l], az));
                const float cz = __fsub_rn(__fmul_rn(rx[l], ay), __fmul_rn(ry[l], ax));
                float sn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz)));
                float cs = __fadd_rn(__fadd_rn(__fmul_rn(rx[l], ax), __fmul_rn(ry[l], ay)), __fmul_rn(rz[l], az));
                sn = fmaxf(sn, 1e-8f);                 // :339
                cs = fminf(fmaxf(cs, -1.0f), 1.0f);    // :340: -1+1e-8 and 1-1e-8 are -1 and 1 in fp32
                arow[(size_t)j * k + l] = __fmul_rn(atan2f(sn, cs), factor_a);
            }
        }
    }
}

}  // namespace pem

using namespace pem;

extern "C" int pem_geo_embedding_indices(const float *points, int B, int N, int angle_k, float sigma_d,
                                         float sigma_a, float *d_indices, float *a_indices,
                                         pem_stream_t stream) {
    PEM_REQUIRE_PTR(points);
    PEM_REQUIRE_PTR(d_indices);
    PEM_REQUIRE_PTR(a_indices);
    PEM_REQUIRE(B > 0 && N > 0 && angle_k > 0 && sigma_d > 0.f && sigma_a > 0.f && B <= 65535, PEM_E_SHAPE);
    PEM_REQUIRE(angle_k <= kGeoMaxK && angle_k < N && N <= 8192, PEM_E_UNSUPPORTED);  // 4 floats of smem per point
    const size_t smem = (size_t)4 * N * sizeof(float);
    static PemSmemCache cached(40 * 1024);
    if (int rc = pem_ensure_dyn_smem(geo_embedding_kernel, smem, &cached)) return rc;
    const float factor_a = (float)(180.0 / ((double)sigma_a * 3.14159265358979323846));  // transformer.py:294, a Python float
    geo_embedding_kernel<<<dim3(N, B), kGeoThreads, smem, pem_cu_stream(stream)>>>(points, N, angle_k, sigma_d,
                                                                                 factor_a, d_indices, a_indices);
    return pem_launch_status();
}
