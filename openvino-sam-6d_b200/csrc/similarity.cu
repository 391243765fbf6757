// compute_feature_similarity ('cosine', utils/model_utils.py:131-153 of the reference):
// atten[b,i,j] = <f1[b,i]/max(|f1[b,i]|,1e-12), f2[b,j]/max(|f2[b,j]|,1e-12)> / temp.
//
// This file holds the fp32-exact path: a register-tiled FFMA contraction (both operands
// K-major, "NT"), row norms applied in the epilogue.  It is the path used for the coarse
// stage (197x197x256, feeds the inverse-CDF sampler, so it must be fp32-accurate) and the
// fp32 reference mode of the fine stage; the tcgen05 split-bf16 fine kernel lives in
// fine_tc.cu.
#include "pem_common.cuh"
#include "tc_ptx.cuh"  // packed fp32 pairs

namespace pem {

// 1 / max(||row||_2, 1e-12): one warp per row (F.normalize, eps = 1e-12); both feature sets in one
// launch (rows [0, rows1) of f1, then rows2 rows of f2; inv holds them back to back).
__global__ void __launch_bounds__(256)
row_inv_norm_kernel(const float *__restrict__ f1, int rows1, const float *__restrict__ f2, int rows2, int C,
                    float *__restrict__ inv) {
    const int row = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (row >= rows1 + rows2) return;
    const float *p = row < rows1 ? f1 + (size_t)row * C : f2 + (size_t)(row - rows1) * C;
    float s = 0.f;
    for (int c = lane_id(); c < C; c += 32) {
        const float v = __ldg(p + c);
        s = fmaf(v, v, s);
    }
    s = warp_sum(s);
    if (lane_id() == 0) inv[row] = 1.0f / fmaxf(sqrtf(s), 1e-12f);
}

// C[b] (N1 x N2) = A[b] (N1 x K) * B[b]^T (N2 x K); 256 threads, (BM/TM) x (BN/TN) = 16 x 16.
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256)
similarity_nt_kernel(const float *__restrict__ A, const float *__restrict__ Bm, int N1, int N2,
                     int K, const float *__restrict__ invA, const float *__restrict__ invB,
                     float temp, float *__restrict__ out) {
    constexpr int BK = 16;
    static_assert(BM / TM == 16 && BN / TN == 16, "256-thread layout");
    __shared__ float As[2][BK][BM + 4];
    __shared__ float Bs[2][BK][BN + 4];
    const int b = blockIdx.z;
    const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    const float *Ab = A + (size_t)b * N1 * K;
    const float *Bb = Bm + (size_t)b * N2 * K;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

    float acc[TM][TN];
#pragma unroll
    for (int m = 0; m < TM; ++m)
#pragma unroll
        for (int n = 0; n < TN; ++n) acc[m][n] = 0.f;

    // each thread moves LA (LB) float4 per K-step: row = (tid + l*256) / 4, kq = (tid + l*256) % 4
    constexpr int LA = BM * BK / 4 / 256, LB = BN * BK / 4 / 256;
    float4 ra[LA], rb[LB];
    auto g_load = [&](int k0) {
#pragma unroll
        for (int l = 0; l < LA; ++l) {
            const int q = tid + l * 256, r = q >> 2, kq = (q & 3) * 4;
            const int gi = i0 + r, gk = k0 + kq;
            ra[l] = (gi < N1 && gk + 3 < K)
                        ? __ldg(reinterpret_cast<const float4 *>(Ab + (size_t)gi * K + gk))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
            if (gi < N1 && gk < K && gk + 3 >= K) {  // ragged K tail
                const float *p = Ab + (size_t)gi * K;
                ra[l].x = p[gk];
                ra[l].y = gk + 1 < K ? p[gk + 1] : 0.f;
                ra[l].z = gk + 2 < K ? p[gk + 2] : 0.f;
                ra[l].w = 0.f;
            }
        }
#pragma unroll
        for (int l = 0; l < LB; ++l) {
            const int q = tid + l * 256, r = q >> 2, kq = (q & 3) * 4;
            const int gj = j0 + r, gk = k0 + kq;
            rb[l] = (gj < N2 && gk + 3 < K)
                        ? __ldg(reinterpret_cast<const float4 *>(Bb + (size_t)gj * K + gk))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
            if (gj < N2 && gk < K && gk + 3 >= K) {
                const float *p = Bb + (size_t)gj * K;
                rb[l].x = p[gk];
                rb[l].y = gk + 1 < K ? p[gk + 1] : 0.f;
                rb[l].z = gk + 2 < K ? p[gk + 2] : 0.f;
                rb[l].w = 0.f;
            }
        }
    };
    auto s_store = [&](int buf) {
#pragma unroll
        for (int l = 0; l < LA; ++l) {
            const int q = tid + l * 256, r = q >> 2, kq = (q & 3) * 4;
            As[buf][kq + 0][r] = ra[l].x; As[buf][kq + 1][r] = ra[l].y;
            As[buf][kq + 2][r] = ra[l].z; As[buf][kq + 3][r] = ra[l].w;
        }
#pragma unroll
        for (int l = 0; l < LB; ++l) {
            const int q = tid + l * 256, r = q >> 2, kq = (q & 3) * 4;
            Bs[buf][kq + 0][r] = rb[l].x; Bs[buf][kq + 1][r] = rb[l].y;
            Bs[buf][kq + 2][r] = rb[l].z; Bs[buf][kq + 3][r] = rb[l].w;
        }
    };

    const int nk = (K + BK - 1) / BK;
    g_load(0);
    s_store(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) g_load((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], bb[TN];
#pragma unroll
            for (int m = 0; m < TM; m += 4) {
                const float4 v = *reinterpret_cast<const float4 *>(&As[buf][k][ty * TM + m]);
                a[m] = v.x; a[m + 1] = v.y; a[m + 2] = v.z; a[m + 3] = v.w;
            }
#pragma unroll
            for (int n = 0; n < TN; n += 4) {
                const float4 v = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * TN + n]);
                bb[n] = v.x; bb[n + 1] = v.y; bb[n + 2] = v.z; bb[n + 3] = v.w;
            }
            // two columns per instruction (FFMA2: the same fp32 fma per element, half the issue slots)
#pragma unroll
            for (int m = 0; m < TM; ++m) {
                const uint64_t am = tc::pack2(a[m], a[m]);
#pragma unroll
                for (int n = 0; n < TN; n += 2)
                    tc::unpack2(tc::fma2(am, tc::pack2(bb[n], bb[n + 1]), tc::pack2(acc[m][n], acc[m][n + 1])),
                                acc[m][n], acc[m][n + 1]);
            }
        }
        if (kt + 1 < nk) {
            s_store(buf ^ 1);
            __syncthreads();
        }
    }

    float *ob = out + (size_t)b * N1 * N2;
#pragma unroll
    for (int m = 0; m < TM; ++m) {
        const int gi = i0 + ty * TM + m;
        if (gi >= N1) continue;
        const float sa = invA ? __ldg(invA + (size_t)b * N1 + gi) : 1.0f;
#pragma unroll
        for (int n = 0; n < TN; ++n) {
            const int gj = j0 + tx * TN + n;
            if (gj >= N2) continue;
            const float sb = invB ? __ldg(invB + (size_t)b * N2 + gj) : 1.0f;
            ob[(size_t)gi * N2 + gj] = __fdiv_rn(acc[m][n] * sa * sb, temp);
        }
    }
}

}  // namespace pem

using namespace pem;

extern "C" size_t pem_similarity_workspace_bytes(int B, int N1, int N2) {
    if (B <= 0 || N1 <= 0 || N2 <= 0) return 0;
    return (size_t)B * ((size_t)N1 + N2) * sizeof(float);
}

extern "C" int pem_feature_similarity(const float *f1, const float *f2, int B, int N1, int N2, int C,
                                      float temp, int normalize, float *atten, void *workspace,
                                      size_t workspace_bytes, pem_stream_t stream) {
    PEM_REQUIRE_PTR(f1);
    PEM_REQUIRE_PTR(f2);
    PEM_REQUIRE_PTR(atten);
    PEM_REQUIRE(B > 0 && N1 > 0 && N2 > 0 && C > 0 && B <= 65535, PEM_E_SHAPE);
    PEM_REQUIRE(C % 4 == 0 && ((uintptr_t)f1 % 16 == 0) && ((uintptr_t)f2 % 16 == 0),
                PEM_E_UNSUPPORTED);
    cudaStream_t st = pem_cu_stream(stream);
    float *inv1 = nullptr, *inv2 = nullptr;
    if (normalize) {
        PEM_REQUIRE(workspace != nullptr &&
                        workspace_bytes >= pem_similarity_workspace_bytes(B, N1, N2),
                    PEM_E_WORKSPACE);
        inv1 = (float *)workspace;
        inv2 = inv1 + (size_t)B * N1;
        row_inv_norm_kernel<<<(B * (N1 + N2) + 7) / 8, 256, 0, st>>>(f1, B * N1, f2, B * N2, C, inv1);
    }
    if (N1 <= 512 && N2 <= 512) {
        dim3 grid((N2 + 63) / 64, (N1 + 63) / 64, B);
        similarity_nt_kernel<64, 64, 4, 4><<<grid, 256, 0, st>>>(f1, f2, N1, N2, C, inv1, inv2,
                                                                 temp, atten);
    } else {
        dim3 grid((N2 + 127) / 128, (N1 + 127) / 128, B);
        similarity_nt_kernel<128, 128, 8, 8><<<grid, 256, 0, st>>>(f1, f2, N1, N2, C, inv1, inv2,
                                                                   temp, atten);
    }
    return pem_launch_status();
}
