/*
 * oracle/pointnet2_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the reference's CPU pointnet2 ops.  The arithmetic
 * (operation order, float-vs-double promotion, comparison strictness, tie
 * breaking) follows the reference's CPU loops exactly; it is written
 * independently (different loop structure / naming) and is pinned against the
 * unmodified reference build (oracle/_ref) by tests/test_oracle_cpu.py and
 * against tests/golden/*.npz.
 *
 * Reference (paths relative to /root/reference/SAM-6D/Pose_Estimation_Model/):
 *   furthest point sampling  model/pointnet2/_ext_src/src/sampling.cpp:76-118
 *   gather                   model/pointnet2/_ext_src/src/sampling.cpp:23-44
 *   ball query               model/pointnet2/_ext_src/src/ball_query.cpp:16-62
 *   grouping                 model/pointnet2/_ext_src/src/group_points.cpp:20-45
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/Makefile).  No FMA
 * contraction is allowed: the reference build (-O3, x86-64 baseline) contains
 * none, and bit-exact indices depend on it.
 */
#include <float.h>
#include <stddef.h>
#include <stdint.h>

/* squared distance in the reference's evaluation order: ((dx*dx)+(dy*dy))+(dz*dz) */
static float sqdist3(float ax, float ay, float az, float bx, float by, float bz)
{
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    float s = dx * dx;
    s = s + dy * dy;
    s = s + dz * dz;
    return s;
}

/*
 * FPS, CPU semantics (sampling.cpp:76-118):
 *  - idx[0] = 0; running distance initialised to FLT_MAX (sampling.cpp:86)
 *  - a point whose squared norm, promoted to double, is <= 1e-3 is skipped:
 *    never updated, never selectable (sampling.cpp:101-103; `mag <= 1e-3`
 *    compares a float with a double literal)
 *  - running distance = std::min(d, temp) = (temp < d) ? temp : d
 *  - strict `>` against best (initialised to -FLT_MAX) in ascending k, so
 *    ties resolve to the lowest index; if nothing qualifies the result is 0.
 * skip_near_origin = 0 reproduces the reference's CUDA variant's rule
 * (sampling_gpu.cu:106 has the skip commented out) with CPU tie-breaking.
 */
void oracle_fps(const float *xyz, int B, int N, int M, int skip_near_origin,
                float *scratch /* B*N floats */, int32_t *idx /* B*M */)
{
    if (M <= 0) return;
    for (int b = 0; b < B; ++b) {
        const float *p = xyz + (size_t)b * N * 3;
        float *run = scratch + (size_t)b * N;
        int32_t *out = idx + (size_t)b * M;
        for (int k = 0; k < N; ++k) run[k] = FLT_MAX;
        int32_t cur = 0;
        out[0] = 0;
        for (int j = 1; j < M; ++j) {
            const float cx = p[3 * cur], cy = p[3 * cur + 1], cz = p[3 * cur + 2];
            float best = -FLT_MAX;
            int32_t arg = 0;
            for (int k = 0; k < N; ++k) {
                const float x = p[3 * k], y = p[3 * k + 1], z = p[3 * k + 2];
                if (skip_near_origin) {
                    float mag = x * x;
                    mag = mag + y * y;
                    mag = mag + z * z;
                    if ((double)mag <= 1e-3) continue;
                }
                const float d = sqdist3(x, y, z, cx, cy, cz);
                const float t = run[k];
                const float m = (t < d) ? t : d;
                run[k] = m;
                if (m > best) { best = m; arg = k; }
            }
            cur = arg;
            out[j] = arg;
        }
    }
}

/* gather (sampling.cpp:23-44): out[b,c,j] = feat[b,c,idx[b,j]], 0 for an index outside [0,N). */
void oracle_gather(const float *feat, const int32_t *idx, int B, int C, int N, int M, float *out)
{
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            const float *row = feat + ((size_t)b * C + c) * N;
            float *o = out + ((size_t)b * C + c) * M;
            for (int j = 0; j < M; ++j) {
                const int32_t a = idx[(size_t)b * M + j];
                o[j] = (a >= 0 && a < N) ? row[a] : 0.0f;
            }
        }
}

/*
 * ball query (ball_query.cpp:16-62): for each centre, the first `nsample`
 * points in index order with d2 < radius*radius (fp32 product, strict <);
 * the first hit back-fills every slot; a centre with no hit keeps the
 * caller's zero initialisation (ball_query.cpp:75-77).  d2 is evaluated as
 * (centre - point) squared, summed x,y,z in order.
 */
void oracle_ball_query(const float *centres, const float *xyz, int B, int M, int N,
                       float radius, int nsample, int32_t *idx /* B*M*nsample, pre-zeroed */)
{
    const float r2 = radius * radius;
    for (int b = 0; b < B; ++b) {
        const float *p = xyz + (size_t)b * N * 3;
        const float *q = centres + (size_t)b * M * 3;
        for (int j = 0; j < M; ++j) {
            int32_t *row = idx + ((size_t)b * M + j) * nsample;
            int found = 0;
            for (int k = 0; k < N && found < nsample; ++k) {
                const float d2 = sqdist3(q[3 * j], q[3 * j + 1], q[3 * j + 2],
                                         p[3 * k], p[3 * k + 1], p[3 * k + 2]);
                if (!(d2 < r2)) continue;
                if (found == 0)
                    for (int l = 0; l < nsample; ++l) row[l] = k;
                row[found++] = k;
            }
        }
    }
}

/* grouping (group_points.cpp:20-45): out[b,c,j,k] = feat[b,c,idx[b,j,k]], 0 for invalid index. */
void oracle_group(const float *feat, const int32_t *idx, int B, int C, int N, int M, int ns,
                  float *out)
{
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            const float *row = feat + ((size_t)b * C + c) * N;
            for (size_t e = 0; e < (size_t)M * ns; ++e) {
                const int32_t a = idx[(size_t)b * M * ns + e];
                out[((size_t)b * C + c) * M * ns + e] = (a >= 0 && a < N) ? row[a] : 0.0f;
            }
        }
}
