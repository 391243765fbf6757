// Register-resident 3x3 SVD / Kabsch rotation for the Procrustes solves of the PEM hot
// path (reference: utils/model_utils.py:343-436, `weighted_procrustes`).
//
// The reference computes U,S,V = torch.svd(H) (LAPACK gesdd) and
//     R = V diag(1,1,sign(det(V U^T))) U^T                      (model_utils.py:415-420)
// For orthonormal U,V this equals
//     R = v1 u1^T + v2 u2^T + (v1 x v2)(u1 x u2)^T
// because u1 x u2 = det(U) u3 and v1 x v2 = det(V) v3, so the sign/ordering ambiguities of
// the third singular pair cancel exactly as the det-correction does.  Only the two leading
// singular pairs are needed, which also makes the rank-2 case (every 3-point hypothesis:
// three centred points are coplanar) well conditioned.
//
// The decomposition is a one-sided (Hestenes) Jacobi iteration on the columns of H,
// carried out in fp64: H itself is formed in fp32 like the reference, and solving that H
// "exactly" halves the discrepancy against LAPACK's fp32 answer.  Degenerate inputs:
// H = 0 gives R = I (LAPACK returns U = V = I there); rank-1 H has a one-parameter family
// of optimal rotations and LAPACK's choice is not reproducible -- we return a valid
// member of the family (proper rotation, orthonormal).
#pragma once

namespace pem {

__device__ __forceinline__ void jacobi_pair(double G[3][3], double V[3][3], int p, int q) {
    double a = 0.0, b = 0.0, g = 0.0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        a = fma(G[r][p], G[r][p], a);
        b = fma(G[r][q], G[r][q], b);
        g = fma(G[r][p], G[r][q], g);
    }
    if (g * g <= 1e-30 * a * b || g == 0.0) return;  // already orthogonal
    const double zeta = (b - a) / (2.0 * g);
    const double tt = copysign(1.0, zeta) / (fabs(zeta) + sqrt(fma(zeta, zeta, 1.0)));
    const double c = rsqrt(fma(tt, tt, 1.0)), s = c * tt;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double gp = G[r][p], gq = G[r][q];
        G[r][p] = c * gp - s * gq;
        G[r][q] = s * gp + c * gq;
        const double vp = V[r][p], vq = V[r][q];
        V[r][p] = c * vp - s * vq;
        V[r][q] = s * vp + c * vq;
    }
}

__device__ __forceinline__ void cross3(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

// R (row-major 3x3) maximising tr(R H) ... i.e. the reference's V diag(1,1,d) U^T for
// H = U S V^T.  H is row-major: H[a][c].
__device__ __forceinline__ void kabsch_rotation(const float H[3][3], float R[3][3]) {
    double G[3][3], V[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            G[r][c] = (double)H[r][c];
            V[r][c] = (r == c) ? 1.0 : 0.0;
        }
#pragma unroll 1
    for (int sweep = 0; sweep < 6; ++sweep) {
        jacobi_pair(G, V, 0, 1);
        jacobi_pair(G, V, 0, 2);
        jacobi_pair(G, V, 1, 2);
    }
    // column norms^2 = squared singular values; sort columns descending with static
    // indexing (compare-exchange network) so everything stays in registers
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
    constexpr int i1 = 0, i2 = 1;

    if (!(n[i1] > 0.0)) {  // H == 0 (or NaN): identity, as LAPACK's U = V = I
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) R[r][c] = (r == c) ? 1.f : 0.f;
        return;
    }
    double u1[3], u2[3], v1[3], v2[3];
    const double inv1 = rsqrt(n[i1]);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        u1[r] = G[r][i1] * inv1;
        v1[r] = V[r][i1];
        v2[r] = V[r][i2];
    }
    // u2: Gram-Schmidt against u1 (exact orthogonality even when sigma2 is small)
    double g2[3] = {G[0][i2], G[1][i2], G[2][i2]};
    const double pr = g2[0] * u1[0] + g2[1] * u1[1] + g2[2] * u1[2];
#pragma unroll
    for (int r = 0; r < 3; ++r) g2[r] -= pr * u1[r];
    double l2 = g2[0] * g2[0] + g2[1] * g2[1] + g2[2] * g2[2];
    if (l2 > 1e-28 * n[i1]) {
        const double inv2 = rsqrt(l2);
#pragma unroll
        for (int r = 0; r < 3; ++r) u2[r] = g2[r] * inv2;
    } else {
        // rank 1: any unit vector orthogonal to u1 (cross with the smallest-component axis)
        const double a0 = fabs(u1[0]), a1 = fabs(u1[1]), a2 = fabs(u1[2]);
        double e[3];
        e[0] = (a0 <= a1 && a0 <= a2) ? 1.0 : 0.0;
        e[1] = (e[0] == 0.0 && a1 <= a2) ? 1.0 : 0.0;
        e[2] = (e[0] == 0.0 && e[1] == 0.0) ? 1.0 : 0.0;
        cross3(u1, e, u2);
        const double inv2 = rsqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
#pragma unroll
        for (int r = 0; r < 3; ++r) u2[r] *= inv2;
    }
    double u3[3], v3[3];
    cross3(u1, u2, u3);
    cross3(v1, v2, v3);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            R[r][c] = (float)(v1[r] * u1[c] + v2[r] * u2[c] + v3[r] * u3[c]);
}

}  // namespace pem
