// Host-side instance preparation of the reference's inference script, on the device (SURVEY.md
// section 8f, rank 4): depth back-projection (utils/data_utils.py:92-110), and per detection the
// geometric part of get_test_data's loop (run_inference_custom_pytorch.py:303-351): mask & depth > 0,
// the square box of get_bbox (data_utils.py:125-160), the crop-flat indices of the masked pixels, the
// distance filter around their mean, and -- after the HOST has drawn the sample with numpy's
// generator, which is what keeps the result identical to the reference's for a given seed -- the
// gather of the drawn points and get_resize_rgb_choose (data_utils.py:112-123).
//
// Arithmetic.  The reference lines mix float32 arrays with the float64 camera scalars of the json
// file.  Under its pinned numpy 1.26 that stays float32 (value-based casting); under NumPy >= 2 the
// same lines promote to float64 and round once when the cloud becomes a FloatTensor.  Both are
// implemented (precision 0 = f32, 1 = f64), one IEEE operation per numpy operation, no contraction.
// np.mean over axis 0 of an (n,3) array adds row after row, so the centre is a SEQUENTIAL sum: three
// threads add the compacted points of a chunk in order while the CTA stages the next chunk.
//
// One CTA per detection; the full-image point cloud is never materialised per detection (the
// reference copies the whole (H,W,3) cloud once per detection): coordinates are recomputed from the
// depth image wherever they are needed.
#include "pem_common.cuh"

namespace pem {

constexpr int kPrepThreads = 1024;
constexpr int kInfoInts = 8;  // valid | n_mask | n_valid | y1 | y2 | x1 | x2 | pad

struct Camera { double fx, fy, cx, cy; };

template <typename T> struct Arith;
template <> struct Arith<float> {
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
};
template <> struct Arith<double> {
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
};

// pt0 = (x - cx) * z / fx ; pt1 = (y - cy) * z / fy ; pt2 = z
template <typename T>
__device__ __forceinline__ void backproject(int row, int col, float z, const Camera &k, T &x, T &y, T &zz) {
    using A = Arith<T>;
    x = A::div(A::mul(A::sub((T)(float)col, (T)k.cx), (T)z), (T)k.fx);
    y = A::div(A::mul(A::sub((T)(float)row, (T)k.cy), (T)z), (T)k.fy);
    zz = (T)z;
}

__global__ void __launch_bounds__(256)
depth_to_metres_kernel(const uint16_t *__restrict__ raw, int n, float scale, float *__restrict__ out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) out[i] = __fdiv_rn(__fmul_rn((float)raw[i], scale), 1000.0f);
}

template <typename T>
__global__ void __launch_bounds__(256)
backproject_kernel(const float *__restrict__ depth, int H, int W, Camera k, float *__restrict__ cloud) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= H * W) return;
    const int row = i / W, col = i - row * W;
    T x, y, z;
    backproject<T>(row, col, __ldg(depth + i), k, x, y, z);
    cloud[3 * (size_t)i] = (float)x;
    cloud[3 * (size_t)i + 1] = (float)y;
    cloud[3 * (size_t)i + 2] = (float)z;
}

// position of this thread's element among the CTA's set predicates (in thread order) and their total
__device__ __forceinline__ int cta_rank(bool pred, int *s_warp /* [33] */, int &total) {
    const unsigned bal = __ballot_sync(kFull, pred);
    const int w = warp_id(), l = lane_id();
    __syncthreads();  // s_warp may still be read from the previous call
    if (l == 0) s_warp[w] = __popc(bal);
    __syncthreads();
    if (w == 0) {
        const int v = s_warp[l];
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, inc, o);
            if (l >= o) inc += t;
        }
        s_warp[l] = inc - v;
        if (l == 31) s_warp[32] = inc;
    }
    __syncthreads();
    total = s_warp[32];
    return s_warp[w] + __popc(bal & ((1u << l) - 1u));
}

template <typename T>
__global__ void __launch_bounds__(kPrepThreads, 1)
instance_candidates_kernel(const float *__restrict__ depth, const uint8_t *__restrict__ masks, int H, int W,
                           Camera cam, T dist_thresh, int32_t *__restrict__ info,
                           int32_t *__restrict__ scratch /* (D, H*W) */) {
    using A = Arith<T>;
    __shared__ int s_warp[33];
    __shared__ int s_red[5][32];
    __shared__ int s_box[4];
    __shared__ T s_pt[3][kPrepThreads];
    __shared__ T s_acc[3];
    const int d = blockIdx.x, tid = threadIdx.x, HW = H * W;
    const uint8_t *mask = masks + (size_t)d * HW;
    int32_t *choose = scratch + (size_t)d * HW;
    int32_t *inf = info + (size_t)d * kInfoInts;

    // A. mask & depth > 0: pixel count and the tight box
    int cnt = 0, rmin = H, rmax = -1, cmin = W, cmax = -1;
    for (int i = tid; i < HW; i += kPrepThreads) {
        if (mask[i] != 0 && __ldg(depth + i) > 0.f) {
            const int r = i / W, c = i - r * W;
            ++cnt;
            rmin = min(rmin, r); rmax = max(rmax, r);
            cmin = min(cmin, c); cmax = max(cmax, c);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(kFull, cnt, o);
        rmin = min(rmin, __shfl_xor_sync(kFull, rmin, o));
        rmax = max(rmax, __shfl_xor_sync(kFull, rmax, o));
        cmin = min(cmin, __shfl_xor_sync(kFull, cmin, o));
        cmax = max(cmax, __shfl_xor_sync(kFull, cmax, o));
    }
    if (lane_id() == 0) {
        const int w = warp_id();
        s_red[0][w] = cnt; s_red[1][w] = rmin; s_red[2][w] = rmax; s_red[3][w] = cmin; s_red[4][w] = cmax;
    }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kPrepThreads / 32; ++w) {
            cnt += s_red[0][w];
            rmin = min(rmin, s_red[1][w]); rmax = max(rmax, s_red[2][w]);
            cmin = min(cmin, s_red[3][w]); cmax = max(cmax, s_red[4][w]);
        }
        // get_bbox: square of side min(max(rows, cols), min(H, W)) around the integer centre, pushed
        // back inside the image (all quantities non-negative, so int() is a plain truncation)
        int r0 = rmin, r1 = rmax + 1, c0 = cmin, c1 = cmax + 1;
        if (cnt > 0) {
            const int b = min(max(r1 - r0, c1 - c0), min(H, W));
            const int cr = (r0 + r1) / 2, cc = (c0 + c1) / 2, half = b / 2;
            r0 = cr - half; r1 = cr + half; c0 = cc - half; c1 = cc + half;
            if (r0 < 0) { r1 -= r0; r0 = 0; }
            if (c0 < 0) { c1 -= c0; c0 = 0; }
            if (r1 > H) { r0 -= r1 - H; r1 = H; }
            if (c1 > W) { c0 -= c1 - W; c1 = W; }
        }
        s_box[0] = r0; s_box[1] = r1; s_box[2] = c0; s_box[3] = c1;
        inf[0] = 0; inf[1] = cnt; inf[2] = 0;
        inf[3] = r0; inf[4] = r1; inf[5] = c0; inf[6] = c1; inf[7] = 0;
        s_red[0][0] = cnt;
    }
    if (tid < 3) s_acc[tid] = (T)0;
    __syncthreads();
    if (s_red[0][0] <= 32) return;  // "if np.sum(mask) > 32 ... else continue"
    const int y1 = s_box[0], y2 = s_box[1], x1 = s_box[2], x2 = s_box[3];
    const int cw = x2 - x1, area = (y2 - y1) * cw;

    // B. crop-flat indices of the masked pixels, ascending, and the sequential sum of their points.
    // (the square box can cut pixels of the mask off: the count inside the crop is what matters now)
    int n = 0;
    for (int base = 0; base < area; base += kPrepThreads) {
        const int i = base + tid;
        bool pred = false;
        int row = 0, col = 0;
        float z = 0.f;
        if (i < area) {
            row = y1 + i / cw; col = x1 + i % cw;
            const int pix = row * W + col;
            z = __ldg(depth + pix);
            pred = mask[pix] != 0 && z > 0.f;
        }
        int total;
        const int pos = cta_rank(pred, s_warp, total);
        if (pred) {
            choose[n + pos] = i;
            T x, y, zz;
            backproject<T>(row, col, z, cam, x, y, zz);
            s_pt[0][pos] = x; s_pt[1][pos] = y; s_pt[2][pos] = zz;
        }
        __syncthreads();
        if (tid < 3) {
            T acc = s_acc[tid];
            for (int j = 0; j < total; ++j) acc = A::add(acc, s_pt[tid][j]);
            s_acc[tid] = acc;
        }
        n += total;
        // (cta_rank's leading barrier orders these reads before the next chunk's writes)
    }
    __syncthreads();
    if (n == 0) return;  // cannot happen when the count above exceeded 32 unless the box cut everything
    const T cx = A::div(s_acc[0], (T)n), cy = A::div(s_acc[1], (T)n), cz = A::div(s_acc[2], (T)n);

    // C. keep the points closer than dist_thresh to the mean; compaction in place (a survivor's new
    // position never exceeds its old one, and a chunk is read completely before it is written)
    int m = 0;
    for (int base = 0; base < n; base += kPrepThreads) {
        const int j = base + tid;
        bool pred = false;
        int i = 0;
        if (j < n) {
            i = choose[j];
            const int row = y1 + i / cw, col = x1 + i % cw;
            T x, y, z;
            backproject<T>(row, col, __ldg(depth + row * W + col), cam, x, y, z);
            const T dx = A::sub(x, cx), dy = A::sub(y, cy), dz = A::sub(z, cz);
            const T s = A::add(A::add(A::mul(dx, dx), A::mul(dy, dy)), A::mul(dz, dz));
            pred = A::sqrt(s) < dist_thresh;
        }
        int total;
        const int pos = cta_rank(pred, s_warp, total);
        if (pred) choose[m + pos] = i;
        m += total;
    }
    if (tid == 0) {
        inf[1] = n;
        inf[2] = m;
        inf[0] = m >= 4 ? 1 : 0;  // "if np.sum(flag) < 4: continue"
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
instance_sample_kernel(const float *__restrict__ depth, int H, int W, Camera cam, const int32_t *__restrict__ info,
                       const int32_t *__restrict__ scratch, const int32_t *__restrict__ det_ids,
                       const long long *__restrict__ pick, int n_sample, int img_size,
                       float *__restrict__ pts, long long *__restrict__ rgb_choose,
                       long long *__restrict__ choose_out, int D) {
    const int slot = blockIdx.y, j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n_sample) return;
    const int d = det_ids[slot];
    const bool det_ok = d >= 0 && d < D;  // a detection id outside [0,D): flagged like a bad draw, never read
    const int32_t *inf = info + (size_t)(det_ok ? d : 0) * kInfoInts;
    const int m = inf[2], y1 = inf[3], y2 = inf[4], x1 = inf[5], x2 = inf[6];
    const size_t o = (size_t)slot * n_sample + j;
    long long p = pick[o];
    if (!det_ok || p < 0 || p >= m) {  // a draw outside the candidate list: flagged, never read
        pts[3 * o] = pts[3 * o + 1] = pts[3 * o + 2] = 0.f;
        rgb_choose[o] = -1;
        if (choose_out) choose_out[o] = -1;
        return;
    }
    const int cw = x2 - x1, ch = y2 - y1;
    const int i = scratch[(size_t)d * H * W + p];
    const int row = i / cw, col = i % cw;
    T x, y, z;
    backproject<T>(y1 + row, x1 + col, __ldg(depth + (y1 + row) * W + x1 + col), cam, x, y, z);
    pts[3 * o] = (float)x;
    pts[3 * o + 1] = (float)y;
    pts[3 * o + 2] = (float)z;
    // get_resize_rgb_choose: floor(row * (img_size / crop_h)) * img_size + floor(col * (img_size / crop_w))
    const double rh = __ddiv_rn((double)img_size, (double)ch), rw = __ddiv_rn((double)img_size, (double)cw);
    const double v = __dadd_rn(__dmul_rn(floor(__dmul_rn((double)row, rh)), (double)img_size),
                               floor(__dmul_rn((double)col, rw)));
    rgb_choose[o] = (long long)v;
    if (choose_out) choose_out[o] = i;
}

}  // namespace pem

using namespace pem;

extern "C" int pem_depth_to_metres(const uint16_t *raw, int H, int W, float depth_scale, float *depth_m,
                                   pem_stream_t stream) {
    PEM_REQUIRE_PTR(raw);
    PEM_REQUIRE_PTR(depth_m);
    PEM_REQUIRE(H > 0 && W > 0 && (long long)H * W < (1ll << 30), PEM_E_SHAPE);
    depth_to_metres_kernel<<<(H * W + 255) / 256, 256, 0, pem_cu_stream(stream)>>>(raw, H * W, depth_scale, depth_m);
    return pem_launch_status();
}

extern "C" int pem_backproject_depth(const float *depth_m, int H, int W, double fx, double fy, double cx,
                                     double cy, int precision, float *cloud, pem_stream_t stream) {
    PEM_REQUIRE_PTR(depth_m);
    PEM_REQUIRE_PTR(cloud);
    PEM_REQUIRE(H > 0 && W > 0 && (long long)H * W < (1ll << 30), PEM_E_SHAPE);
    PEM_REQUIRE(precision == 0 || precision == 1, PEM_E_UNSUPPORTED);
    const Camera k{fx, fy, cx, cy};
    cudaStream_t st = pem_cu_stream(stream);
    if (precision == 0)
        backproject_kernel<float><<<(H * W + 255) / 256, 256, 0, st>>>(depth_m, H, W, k, cloud);
    else
        backproject_kernel<double><<<(H * W + 255) / 256, 256, 0, st>>>(depth_m, H, W, k, cloud);
    return pem_launch_status();
}

extern "C" size_t pem_instance_scratch_bytes(int D, int H, int W) {
    if (D <= 0 || H <= 0 || W <= 0) return 0;
    return (size_t)D * H * W * sizeof(int32_t);
}

extern "C" int pem_instance_candidates(const float *depth_m, const uint8_t *masks, int D, int H, int W,
                                       double fx, double fy, double cx, double cy, int precision,
                                       double dist_thresh, int32_t *info, void *scratch, size_t scratch_bytes,
                                       pem_stream_t stream) {
    PEM_REQUIRE_PTR(depth_m);
    PEM_REQUIRE_PTR(masks);
    PEM_REQUIRE_PTR(info);
    PEM_REQUIRE(D > 0 && D <= 65535 && H > 0 && W > 0 && (long long)H * W < (1ll << 30), PEM_E_SHAPE);
    PEM_REQUIRE(precision == 0 || precision == 1, PEM_E_UNSUPPORTED);
    PEM_REQUIRE(scratch != nullptr && scratch_bytes >= pem_instance_scratch_bytes(D, H, W), PEM_E_WORKSPACE);
    const Camera k{fx, fy, cx, cy};
    cudaStream_t st = pem_cu_stream(stream);
    if (precision == 0)
        instance_candidates_kernel<float><<<D, kPrepThreads, 0, st>>>(depth_m, masks, H, W, k, (float)dist_thresh,
                                                                      info, (int32_t *)scratch);
    else
        instance_candidates_kernel<double><<<D, kPrepThreads, 0, st>>>(depth_m, masks, H, W, k, dist_thresh, info,
                                                                       (int32_t *)scratch);
    return pem_launch_status();
}

extern "C" int pem_instance_sample(const float *depth_m, int H, int W, double fx, double fy, double cx, double cy,
                                   int precision, const int32_t *info, const void *scratch, size_t scratch_bytes,
                                   int D, const int32_t *det_ids, int n_kept, const int64_t *pick, int n_sample,
                                   int img_size, float *pts, int64_t *rgb_choose, int64_t *choose,
                                   pem_stream_t stream) {
    PEM_REQUIRE_PTR(depth_m);
    PEM_REQUIRE_PTR(info);
    PEM_REQUIRE_PTR(det_ids);
    PEM_REQUIRE_PTR(pick);
    PEM_REQUIRE_PTR(pts);
    PEM_REQUIRE_PTR(rgb_choose);
    PEM_REQUIRE(D > 0 && H > 0 && W > 0 && n_kept > 0 && n_kept <= 65535 && n_sample > 0 && img_size > 0,
                PEM_E_SHAPE);
    PEM_REQUIRE(precision == 0 || precision == 1, PEM_E_UNSUPPORTED);
    PEM_REQUIRE(scratch != nullptr && scratch_bytes >= pem_instance_scratch_bytes(D, H, W), PEM_E_WORKSPACE);
    const Camera k{fx, fy, cx, cy};
    cudaStream_t st = pem_cu_stream(stream);
    const dim3 grid((n_sample + 255) / 256, n_kept);
    if (precision == 0)
        instance_sample_kernel<float><<<grid, 256, 0, st>>>(depth_m, H, W, k, info, (const int32_t *)scratch, det_ids,
                                                            (const long long *)pick, n_sample, img_size, pts,
                                                            (long long *)rgb_choose, (long long *)choose, D);
    else
        instance_sample_kernel<double><<<grid, 256, 0, st>>>(depth_m, H, W, k, info, (const int32_t *)scratch, det_ids,
                                                             (const long long *)pick, n_sample, img_size, pts,
                                                             (long long *)rgb_choose, (long long *)choose, D);
    return pem_launch_status();
}
