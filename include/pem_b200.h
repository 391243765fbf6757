/*
 * pem_b200.h -- C ABI of libpem_b200.so: the SAM-6D Pose-Estimation-Model geometric
 * hot path as hand-written sm_100a CUDA kernels.
 *
 * This is the drop-in boundary ("B3").  Every entry point
 *   - takes raw DEVICE pointers, explicit sizes and a CUDA stream (passed as void*,
 *     i.e. a cudaStream_t; NULL = the legacy default stream),
 *   - enqueues its kernels on that stream and returns WITHOUT synchronising,
 *   - never allocates: outputs and scratch are caller-owned (see the
 *     *_workspace_bytes queries),
 *   - returns 0 on success, a negative PEM_E_* code for an invalid argument, or a
 *     positive cudaError_t if a launch failed (the reference's CUDA_CHECK_ERRORS
 *     prints and exit(-1)s instead: _ext_src/include/cuda_utils.h:42-51).
 *
 * All floating tensors are contiguous fp32, all index tensors contiguous int32,
 * exactly as the reference's TORCH_CHECKs demand (_ext_src/include/utils.h:20-35).
 *
 * Reference paths below are relative to
 * /root/reference/SAM-6D/Pose_Estimation_Model/ ; MU = utils/model_utils.py.
 */
#ifndef PEM_B200_H
#define PEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PEM_OK 0
#define PEM_E_NULL (-1)      /* required pointer is NULL            */
#define PEM_E_SHAPE (-2)     /* a size is <= 0 or out of range      */
#define PEM_E_UNSUPPORTED (-3)
#define PEM_E_WORKSPACE (-4) /* workspace missing / too small       */

/* FPS flags */
#define PEM_FPS_SKIP_NEAR_ORIGIN 1 /* CPU-reference rule: |p|^2 <= 1e-3 never sampled */

typedef void *pem_stream_t; /* cudaStream_t */

/* Library / build identification. */
int pem_abi_version(void);
const char *pem_build_info(void);
const char *pem_status_string(int status);

/* ---------------------------------------------------------------- pointnet2 ops */

/*
 * Furthest point sampling.  Replaces `furthest_point_sampling`
 * (model/pointnet2/_ext_src/src/sampling.cpp:184-212; CPU loop :76-118, CUDA kernel
 * sampling_gpu.cu:75-178).  xyz (B,N,3) -> idx (B,M) int32, idx[:,0] = 0.
 * Semantics are the reference's CPU loop, bit for bit: unfused fp32 distance
 * ((dx*dx)+(dy*dy))+(dz*dz), lowest index wins ties, and with
 * PEM_FPS_SKIP_NEAR_ORIGIN points with |p|^2 <= 1e-3 (compared in double) are
 * never selected.  flags = 0 drops the near-origin rule (the reference's CUDA variant,
 * sampling_gpu.cu:106) but keeps lowest-index tie breaking.
 * workspace: pem_fps_workspace_bytes(B,N) bytes (0 for N <= 262144 = 16 CTAs x 16384 points, the
 * largest cloud the cluster kernel keeps on chip; may then be NULL).
 */
size_t pem_fps_workspace_bytes(int B, int N);
int pem_fps(const float *xyz, int B, int N, int M, int flags, int32_t *idx, void *workspace,
            size_t workspace_bytes, pem_stream_t stream);
/*
 * The same plus the sampled points themselves, new_xyz (B,M,3) = xyz[b, idx[b,j], :] -- the point half
 * of sample_pts_feats (utils/model_utils.py:70-84: furthest_point_sample + gather_operation on the
 * transposed cloud).  One launch where a cloud lives in one CTA (N <= 2048, the PEM shapes: the sampling
 * thread stores the point it has just chosen); larger clouds run pem_fps and a row gather.
 */
int pem_fps_gather(const float *xyz, int B, int N, int M, int flags, int32_t *idx, float *new_xyz,
                   void *workspace, size_t workspace_bytes, pem_stream_t stream);

/*
 * gather_points (sampling.cpp:120-150; CPU loop :23-44; CUDA sampling_gpu.cu:13-35).
 * feat (B,C,N), idx (B,M) -> out (B,C,M); an index outside [0,N) yields 0.0f (CPU rule).
 */
int pem_gather(const float *feat, const int32_t *idx, int B, int C, int N, int M, float *out,
               pem_stream_t stream);

/*
 * Row gather for point-major tensors: feat (B,N,C) with row stride `row_stride`
 * floats and batch stride `batch_stride` floats, idx (B,M) -> out (B,M,C) with rows
 * contiguous and batch stride `out_batch_stride` floats (lets the caller gather straight
 * into rows 1.. of a background-token-augmented (B,M+1,C) matrix).
 * This is `gather_operation(x.transpose(1,2).contiguous(), idx).transpose(1,2)` of
 * MU:77-80 and transformer.py:667-671 without the two transposes.
 */
int pem_gather_rows(const float *feat, const int32_t *idx, int B, int N, int C, int M,
                    long long batch_stride, long long row_stride, float *out,
                    long long out_batch_stride /* 0 = M*C */, pem_stream_t stream);

/*
 * out[b,i,:] = ((pts[b,i,:] - t[b]) @ R[b]) + eps, or pts + eps when R and t are NULL:
 * `p1_ = (p1 - init_t) @ init_R` (fine_point_matching.py:45) and the `pts + 1e-8` query
 * centres of PositionalEncoding (fine_point_matching.py:117).  pts (B,N,3), R (B,9), t (B,3).
 */
int pem_transform_points(const float *pts, const float *R, const float *t, int B, int N, float eps,
                         float *out, pem_stream_t stream);

/*
 * ball_query (ball_query.cpp:64-94; CPU loop :16-62; CUDA ball_query_gpu.cu:14-59).
 * new_xyz (B,M,3), xyz (B,N,3) -> idx (B,M,nsample): the first nsample points in index
 * order with d2 < radius*radius (fp32, strict); remaining slots repeat the first hit;
 * a centre with no hit gets zeros.  idx is fully written (no pre-zeroing needed).
 * nsample <= 128.
 */
int pem_ball_query(const float *new_xyz, const float *xyz, int B, int M, int N, float radius,
                   int nsample, int32_t *idx, pem_stream_t stream);

/*
 * group_points (group_points.cpp:79-108; CPU loop :20-45; CUDA group_points_gpu.cu:13-44).
 * feat (B,C,N), idx (B,M,ns) -> out (B,C,M,ns); invalid index -> 0.0f.
 */
int pem_group(const float *feat, const int32_t *idx, int B, int C, int N, int M, int ns,
              float *out, pem_stream_t stream);

/*
 * Fused QueryAndGroup.forward with use_xyz=True (pointnet2_utils.py:326-403):
 * ball query, grouping of xyz minus the centre, grouping of `feat`, concatenation.
 * xyz (B,N,3), new_xyz (B,M,3), feat (B,C,N) or NULL (C=0) -> out (B,3+C,M,ns).
 * idx_out (B,M,ns) is optional (NULL to skip).  nsample <= 128.
 */
int pem_query_and_group(const float *xyz, const float *new_xyz, const float *feat, int B, int N,
                        int M, int C, float radius, int nsample, float *out, int32_t *idx_out,
                        pem_stream_t stream);

/*
 * Fused PositionalEncoding grouping (fine_point_matching.py:113-137): both QueryAndGroup scales
 * of PEM on one cloud in a single pass, i.e. for s in {1,2}
 *   out_s = QueryAndGroup(radius_s, nsample_s)(xyz, xyz + centre_eps, xyz^T)   (B,6,N,nsample_s)
 * The squared distance is evaluated once per (centre, point) and compared against both radii;
 * results are bit-identical to two pem_query_and_group calls.  N <= 65535, nsample <= 128.
 */
int pem_pe_group2(const float *xyz, int B, int N, float centre_eps, float radius1, int nsample1,
                  float *out1, float radius2, int nsample2, float *out2, pem_stream_t stream);

/*
 * Tensor-core pre-filter for pem_pe_group2.  The exact ball query pays ~13 instructions for each
 * of the N^2 (centre, point) pairs although only a few per cent lie near the ball, and for the
 * cloud that fine_point_matching.py:113-137 groups AFTER applying the coarse pose that scan sits on
 * the critical path behind the whole coarse stage.
 *   pem_pe_masks          one bit per pair (B,N,ceil(N/32)) uint32: set when the point MAY lie within
 *                         `radius` of the centre -- a split-TF32 tcgen05 GEMM of |p|^2 - 2 c.p with a
 *                         slack that covers its own rounding plus rel_margin * radius^2.  Ball
 *                         membership survives a rigid motion up to rounding, so the masks can be
 *                         computed on the cloud BEFORE it is moved (any time, e.g. a side stream).
 *   pem_pe_group2_masked  pem_pe_group2 on the (moved) cloud visiting only the set bits, each with
 *                         the reference's exact arithmetic (ball_query.cpp:44-61): bit-identical to
 *                         pem_pe_group2 provided the masks were made with radius >= max(radius1,
 *                         radius2) and the motion changes no squared distance by more than
 *                         rel_margin * radius^2 (1e-3 leaves two orders of magnitude over the fp32
 *                         rounding of an orthonormal R).
 * max_ctas > 0 caps the filter's grid (it takes a whole SM per CTA) so that it can share the GPU
 * with a concurrent latency-critical kernel instead of queueing behind it; 0 = one CTA per SM.
 * N <= 4096.
 */
size_t pem_pe_masks_bytes(int B, int N);
int pem_pe_masks(const float *xyz, int B, int N, float centre_eps, float radius, float rel_margin,
                 int max_ctas, void *masks, size_t masks_bytes, pem_stream_t stream);
int pem_pe_group2_masked(const float *xyz, int B, int N, float centre_eps, float radius1,
                         int nsample1, float *out1, float radius2, int nsample2, float *out2,
                         const void *masks, size_t masks_bytes, pem_stream_t stream);

/* ---------------------------------------------------------------- similarity */

/*
 * compute_feature_similarity, 'cosine' (MU:131-153):
 * atten[b,i,j] = <n(f1[b,i]), n(f2[b,j])> / temp, n(x) = x / max(|x|, 1e-12) when
 * normalize != 0.  f1 (B,N1,C), f2 (B,N2,C) -> atten (B,N1,N2).  fp32 FFMA arithmetic
 * (exact-precision path; C % 4 == 0, 16-byte aligned inputs).
 * workspace: pem_similarity_workspace_bytes(B,N1,N2) bytes when normalize != 0.
 */
size_t pem_similarity_workspace_bytes(int B, int N1, int N2);
int pem_feature_similarity(const float *f1, const float *f2, int B, int N1, int N2, int C,
                           float temp, int normalize, float *atten, void *workspace,
                           size_t workspace_bytes, pem_stream_t stream);

/* ---------------------------------------------------------------- coarse stage */

/*
 * Front half of compute_coarse_Rt (MU:230-242): dual softmax, foreground masks,
 * masked score ** 1.5, fp64 prefix sum rounded to fp32 per element, normalised by
 * (last + 1e-8).  atten (B,N1+1,N2+1) -> w1 (B,N1), w2 (B,N2), cdf (B,N1*N2).
 * Requires (N1+1)*(N2+1) <= 50176 (one CTA holds the matrix in shared memory).
 */
int pem_coarse_assign(const float *atten, int B, int N1, int N2, float *w1, float *w2, float *cdf,
                      pem_stream_t stream);

/*
 * Hypothesis generation (MU:248-261, 277-305, 343-436): for every one of the B*P1
 * hypotheses draw 3 correspondences by inverse-CDF lookup of rand (first index with
 * cdf >= r, 0 if none), solve the 3-point Procrustes (weights 1/(3+1e-5)) with a
 * register-resident 3x3 SVD, and evaluate the 3-point residual.
 * cdf (B,N1*N2), rand (B,3*P1), pts1 (B,N1,3), pts2 (B,N2,3) ->
 * R (B,P1,9) row-major, t (B,P1,3), resid (B,P1), corr (B,3*P1) int32 flat
 * correspondence index i1*N2+i2 (optional, may be NULL).
 */
int pem_coarse_hypotheses(const float *cdf, const float *rand, const float *pts1,
                          const float *pts2, int B, int N1, int N2, int P1, float *R, float *t,
                          float *resid, int32_t *corr, pem_stream_t stream);

/*
 * torch.topk(resid, P2, largest=False) as a set (MU:262): the P2 smallest residuals per
 * instance, ties at the threshold resolved to the lowest proposal id, written in
 * ascending proposal-id order.  resid (B,P1) -> sel (B,P2) int32.  NaN sorts last.
 */
int pem_select_smallest(const float *resid, int B, int P1, int P2, int32_t *sel,
                        pem_stream_t stream);

/*
 * Hypothesis scoring (MU:267-271): for each selected hypothesis h transform the N1
 * observed points, q_i = (p1_i - t_h) R_h, take the distance to the nearest of Nm model
 * points (expanded form |q|^2 - 2 q.m + |m|^2 clamped at 0, then sqrt) and form
 * score = sum(w1) / (sum_i w1_i d_i + 1e-8).
 * R (B,P,9), t (B,P,3); sel (B,H) indices into P, or NULL meaning hypothesis h = h
 * (then H <= P); pts1 (B,N1,3), w1 (B,N1), model (B,Nm,3) -> scores (B,H).
 */
int pem_score_hypotheses(const float *R, const float *t, const int32_t *sel, int B, int P, int H,
                         const float *pts1, const float *w1, int N1, const float *model, int Nm,
                         float *scores, pem_stream_t stream);
/* ... with the kernel named by the caller (parity tests compare the two): AUTO = the tensor-core kernel for every
 * shape it takes (<= 4096 model points), else the SIMT kernel; TENSOR returns PEM_E_UNSUPPORTED for other shapes. */
#define PEM_SCORE_AUTO 0
#define PEM_SCORE_SIMT 1
#define PEM_SCORE_TENSOR 2
int pem_score_hypotheses_ex(const float *R, const float *t, const int32_t *sel, int B, int P, int H,
                            const float *pts1, const float *w1, int N1, const float *model, int Nm,
                            float *scores, int kernel, pem_stream_t stream);

/*
 * argmax over scores (MU:272-274), first maximum wins; copies the winner.
 * -> R_out (B,9), t_out (B,3), proposal_id (B) = sel[b,argmax] (or argmax if sel NULL).
 */
int pem_select_best(const float *scores, const int32_t *sel, const float *R, const float *t,
                    int B, int P, int H, float *R_out, float *t_out, int32_t *proposal_id,
                    pem_stream_t stream);

/* ---------------------------------------------------------------- fine stage */

/*
 * compute_fine_Rt (MU:308-341) on a materialised attention matrix.
 * atten (B,N1+1,N2+1), pts1 (B,N1,3), pts2 (B,N2,3), model (B,Nm,3) ->
 * R (B,9), t (B,3), score (B).  Optional diagnostics (may be NULL): w1 (B,N1),
 * w2 (B,N2), rowsum (B,N1), pred (B,N1,3).
 * workspace: pem_fine_workspace_bytes(B,N1,N2).
 */
size_t pem_fine_workspace_bytes(int B, int N1, int N2);
int pem_fine_Rt(const float *atten, const float *pts1, const float *pts2, const float *model,
                int B, int N1, int N2, int Nm, float dis_thres, float *R, float *t, float *score,
                float *w1, float *w2, float *rowsum, float *pred, void *workspace,
                size_t workspace_bytes, pem_stream_t stream);

/*
 * Tensor-core similarity: the same function as pem_feature_similarity computed as a
 * split-bf16 (hi*hi + lo*hi + hi*lo) tcgen05 GEMM with TMA-staged operands and the
 * accumulator in TMEM; logits agree with the fp32 path to ~1e-5.  Requires C % 64 == 0,
 * C <= 256, temp >= 0.025.  f1 (B,N1p,C), f2 (B,N2p,C) -> atten (B,N1p,N2p).
 * workspace: pem_fine_fused_workspace_bytes(B, N1p-1, N2p-1, C).
 */
size_t pem_fine_fused_workspace_bytes(int B, int N1, int N2, int C);
int pem_feature_similarity_tc(const float *f1, const float *f2, int B, int N1p, int N2p, int C,
                              float temp, int normalize, float *atten, void *workspace,
                              size_t workspace_bytes, pem_stream_t stream);

/*
 * Fused fine stage: compute_feature_similarity (MU:131-153, cosine, normalised) followed by
 * compute_fine_Rt (MU:308-341) straight from the features; the (B,N1+1,N2+1) attention
 * matrix is never written to memory.  Two full tcgen05 GEMM passes with reducing epilogues
 * (softmax statistics; foreground labels + weighted sums over every column) and a short ragged
 * pass that takes the background columns' share out again.
 * f1 (B,N1+1,C), f2 (B,N2+1,C) with the background token in row 0; other arguments and
 * outputs as pem_fine_Rt.  Same constraints as pem_feature_similarity_tc.
 */
int pem_fine_fused(const float *f1, const float *f2, const float *pts1, const float *pts2,
                   const float *model, int B, int N1, int N2, int C, int Nm, float temp,
                   float dis_thres, float *R, float *t, float *score, float *w1, float *w2,
                   float *rowsum, float *pred, void *workspace, size_t workspace_bytes,
                   pem_stream_t stream);

/*
 * The same with the GEMM arithmetic chosen by the caller:
 *   PEM_PREC_SPLIT_BF16  every product as bf16 hi*hi + hi*lo + lo*hi (three MMAs): logits to ~1e-5,
 *                        what pem_fine_fused does;
 *   PEM_PREC_FP16        one fp16 x fp16 product with fp32 accumulation.  The operands are unit
 *                        vectors (|x| <= 1, no range problem; 2^-12 relative per element): logits to
 *                        ~1e-3 of their +-1/temp range, final pose within 1e-3 deg / 1e-5 radius
 *                        units of the split form on the parity cases (bar: 0.1 deg, 1e-4), at a third
 *                        of the tensor work and half the operand traffic.  The background row and
 *                        column stay exact fp32 dot products in both modes.
 */
#define PEM_PREC_SPLIT_BF16 0
#define PEM_PREC_FP16 1
/* Options OR-ed into `precision` (same results; the parity tests hold the forms against each other):
 *   PEM_FINE_THREE_PASS         the original three full GEMM passes (statistics, labels, weighted sums) instead of
 *                               two full passes + the ragged background-column correction;
 *   PEM_FINE_STEPS_PER_TASK(n)  n row steps of 256 rows per (instance, strip) task instead of the automatic choice
 *                               (n = 0). */
#define PEM_FINE_THREE_PASS 0x100
#define PEM_FINE_STEPS_PER_TASK(n) (((n) & 0xff) << 16)
int pem_fine_fused_ex(const float *f1, const float *f2, const float *pts1, const float *pts2,
                      const float *model, int B, int N1, int N2, int C, int Nm, float temp,
                      float dis_thres, int precision, float *R, float *t, float *score, float *w1,
                      float *w2, float *rowsum, float *pred, void *workspace, size_t workspace_bytes,
                      pem_stream_t stream);

/*
 * Profiling hook (no reference counterpart): re-run ONE reducing GEMM pass of pem_fine_fused on the
 * operands and constants a previous pem_fine_fused call with the same shapes left in `workspace`:
 * pass 1 softmax statistics, 4 foreground labels + weighted sums over all columns (the two full
 * passes of the default path), 5 the ragged background-column correction launch that follows them,
 * 2 labels only, 3 weighted sums only (the original three-pass path).  Outputs go to workspace scratch only.  bench.py times the dominant kernel with it.
 */
int pem_fine_fused_pass(const float *f1, const float *f2, int B, int N1, int N2, int C, float temp,
                        int pass, void *workspace, size_t workspace_bytes, pem_stream_t stream);
/* ... after a pem_fine_fused_ex call of the same `precision` (pem_fine_fused_pass = SPLIT_BF16).  pass 0 =
 * the operand preparation alone (normalise + convert): it precedes the passes in a step and leaves the
 * operands in L2 for them, which is how bench.py times a pass as it runs inside the step. */
int pem_fine_fused_pass_ex(const float *f1, const float *f2, int B, int N1, int N2, int C, float temp,
                           int pass, int precision, void *workspace, size_t workspace_bytes,
                           pem_stream_t stream);

/*
 * Weighted Procrustes (MU:343-436) for B' problems of n points each, one CTA per
 * problem: src, ref (B',n,3), weights (B',n) or NULL (ones) -> R (B',9), t (B',3).
 * ref ~ R src + t.
 */
int pem_weighted_procrustes(const float *src, const float *ref, const float *weights, int Bp,
                            int n, float weight_thresh, float eps, float *R, float *t,
                            pem_stream_t stream);

/*
 * The reference's operator seam (utils/model_utils.py:457-526): the three autograd `Function` shells whose
 * forward is torch.svd / torch.det / torch.searchsorted and whose `symbolic` emits the custom ONNX ops the
 * OpenVINO plugin fills (ov_pointnet2_op/custom_svd.cpp:103-180 Eigen BDCSVD, custom_det.cpp:52-74,
 * custom_searchsorted.cpp:90-109).  Standalone batched entry points at the shapes of the reference's op
 * scripts (ov_op_test/ov_test_custom_svd.py:31-33: 42000 x 3 x 3); the fused kernels of the hot path carry
 * the same arithmetic inline.
 *   pem_svd3_batched   H (n,3,3) row-major -> U (n,3,3), S (n,3) descending, V (n,3,3) with
 *                      H = U diag(S) V^T (torch.svd's convention: V, not V^T); any of U, S, V may be NULL.
 *                      fp64 one-sided Jacobi per matrix; U, V orthonormal for every input (rank-deficient
 *                      H: missing vectors completed; H = 0 -> U = V = I like LAPACK).  Singular VECTORS are
 *                      unique only up to sign (and rotation inside equal singular values), so parity with
 *                      LAPACK is on S, on U diag(S) V^T, and on R = V diag(1,1,det(V U^T)) U^T.
 *   pem_det3_batched   H (n,3,3) -> det (n): cofactor expansion in fp64, rounded once.
 *   pem_searchsorted   sorted_seq (B,N) non-decreasing, values (B,S) -> out (B,S) int64 = the first e with
 *                      sorted_seq[b,e] >= values[b,s], N when there is none (torch.searchsorted right=False;
 *                      the lower-bound `>=` rule of custom_searchsorted.cpp:95-104).  Bit exact.
 */
int pem_svd3_batched(const float *H, long long n, float *U, float *S, float *V, pem_stream_t stream);
int pem_det3_batched(const float *H, long long n, float *det, pem_stream_t stream);
int pem_searchsorted(const float *sorted_seq, const float *values, int B, int N, int S, int64_t *out,
                     pem_stream_t stream);

/*
 * PositionalEncoding front end (model/fine_point_matching.py:102-144) without the
 * grouped tensor or any activation in memory.
 *   pem_pe_neighbours_masked  the neighbour lists of both PE scales, idx_s (B,N,nsample_s) int32, i.e.
 *                             ball_query(radius_s, nsample_s, xyz, xyz + centre_eps) for s in {1,2}
 *                             (bit-identical), through the candidate masks of pem_pe_masks;
 *   pem_pe_mlp                one PE branch: (neighbour - centre | neighbour) -> SharedMLP [6,32,64,128]
 *                             (1x1 conv without bias, BatchNorm in eval mode, ReLU;
 *                             pointnet2/pytorch_utils.py:25-50,80-134) -> amax over the
 *                             neighbours (fine_point_matching.py:131-133), as three chained split-bf16
 *                             tcgen05 GEMMs per 128 (centre, neighbour) rows.
 * w1 (32,6), w2 (64,32), w3 (128,64) row-major WITH the BatchNorm scale g / sqrt(var + eps) of their
 * output channel multiplied in; shift (224) = b - mean * scale of the three layers' channels back to
 * back; out (B,N,ldo) point-major, this branch in channels [ch_off, ch_off+128) -- two calls fill the
 * (B,N,256) input of the final Conv1d(256, out_dim) = Linear.  nsample in {32, 64, 128}, N <= 4096.
 */
int pem_pe_neighbours_masked(const float *xyz, int B, int N, float centre_eps, float radius1,
                             int nsample1, int32_t *idx1, float radius2, int nsample2, int32_t *idx2,
                             const void *masks, size_t masks_bytes, pem_stream_t stream);
int pem_pe_mlp(const float *xyz, const int32_t *idx, int B, int N, int nsample, float centre_eps,
               const float *w1, const float *w2, const float *w3, const float *shift, float *out,
               int ldo, int ch_off, pem_stream_t stream);

/*
 * GeometricStructureEmbedding.get_embedding_indices (model/transformer.py:304-341):
 * indices of the pair-wise distance embedding and of the triplet-wise angular embedding of the coarse
 * transformer, straight from the sparse points:
 *   d_indices (B,N,N)   = sqrt(pairwise_distance(points, points)) / sigma_d           (:321-322)
 *   a_indices (B,N,N,k) = atan2(|r x a|, r.a) * 180 / (sigma_a * pi) with r = knn_l(i) - p_i, a = p_j - p_i,
 *                         knn = entries 1..k of the k+1 smallest distances of row i, sin clamped at 1e-8
 *                         and the dot product to [-1,1] as the reference does (:325-343)
 * Ties between equal distances go to the lower index.  angle_k <= 8, N <= 8192.
 */
int pem_geo_embedding_indices(const float *points, int B, int N, int angle_k, float sigma_d,
                              float sigma_a, float *d_indices, float *a_indices, pem_stream_t stream);

/*
 * Host-side instance preparation of the inference script, on the device.
 * Replaces the geometric lines of get_test_data (run_inference_custom_pytorch.py:295-351) and the
 * helpers it calls (utils/data_utils.py: get_point_cloud_from_depth :92-110, get_bbox :125-160,
 * get_resize_rgb_choose :112-123).  The RGB crop / cv2.resize and the RLE decoding stay with the caller.
 *
 * precision: 0 = float32 arithmetic (the reference under its pinned numpy 1.26, where a float32 array
 * combined with the float64 camera scalars stays float32), 1 = float64 rounded once at the end (the same
 * lines under NumPy >= 2).  One IEEE operation per numpy operation in either mode.
 *
 * pem_depth_to_metres      depth_m = float32(raw) * depth_scale / 1000                        (:295)
 * pem_backproject_depth    cloud (H*W,3) fp32: ((x - cx) * z / fx, (y - cy) * z / fy, z)        (:296)
 * pem_instance_candidates  per detection d of masks (D,H,W) uint8, one CTA each:              (:311-331)
 *                            mask & depth > 0; dropped unless more than 32 pixels; square box of get_bbox;
 *                            crop-flat indices of the masked pixels in ascending order; their mean as
 *                            numpy takes it (sequential sum over the rows); survivors of
 *                            |p - mean| < dist_thresh (= radius * 1.2 as the caller's numpy evaluates
 *                            it); dropped when fewer than 4 survive.
 *                          info (D,8) int32 = valid | pixels in the crop | survivors | y1 | y2 | x1 | x2 | 0;
 *                          the survivors' indices stay in `scratch` (pem_instance_scratch_bytes).
 * pem_instance_sample      after the HOST has drawn `pick` (n_kept,n_sample) int64 positions into each kept
 *                          detection's survivor list with numpy's generator (np.random.choice in the
 *                          reference's order, :333-338 -- the draw is what makes a seed reproducible, so it
 *                          is not re-implemented): pts (n_kept,n_sample,3) fp32, rgb_choose
 *                          (n_kept,n_sample) int64 = get_resize_rgb_choose of the drawn pixels, choose
 *                          (optional) their crop-flat indices.  det_ids (n_kept) = kept detection numbers.
 *                          A pick outside [0, survivors) yields rgb_choose = -1.
 */
int pem_depth_to_metres(const uint16_t *raw, int H, int W, float depth_scale, float *depth_m,
                        pem_stream_t stream);
int pem_backproject_depth(const float *depth_m, int H, int W, double fx, double fy, double cx, double cy,
                          int precision, float *cloud, pem_stream_t stream);
size_t pem_instance_scratch_bytes(int D, int H, int W);
int pem_instance_candidates(const float *depth_m, const uint8_t *masks, int D, int H, int W, double fx,
                            double fy, double cx, double cy, int precision, double dist_thresh,
                            int32_t *info, void *scratch, size_t scratch_bytes, pem_stream_t stream);
int pem_instance_sample(const float *depth_m, int H, int W, double fx, double fy, double cx, double cy,
                        int precision, const int32_t *info, const void *scratch, size_t scratch_bytes,
                        int D, const int32_t *det_ids, int n_kept, const int64_t *pick, int n_sample,
                        int img_size, float *pts, int64_t *rgb_choose, int64_t *choose,
                        pem_stream_t stream);

/*
 * Pack the per-instance result into the 64-byte record the multi-GPU gather moves
 * (no reference counterpart): records (B,16) fp32 = R (9) | t (3) | pose_score |
 * proposal id | instance id | 0.  proposal_id and instance_id may be NULL.
 */
int pem_pack_records(const float *R, const float *t, const float *score, const int32_t *proposal_id,
                     const float *instance_id, int B, float *records, pem_stream_t stream);

/*
 * Multi-GPU pose-record gather without a collective call ("fused variant"; no reference
 * counterpart): pem_gather_poses packs this rank's n_per_rank 64-byte records (as pem_pack_records) and
 * stores them into the record tables of ALL `world` ranks over NVLink, then raises a per-rank flag;
 * pem_gather_poses_wait waits for every rank's flag of the step and copies the complete
 * (world * n_per_rank, 16) table to `out`.  Both are plain stream work (capturable into the step's CUDA
 * graph).  Tables are double buffered by step parity; every rank must run the same sequence of steps.
 *   setup (once; these call the allocator / CUDA IPC): pem_peer_table_bytes, pem_peer_alloc (cudaMalloc +
 *   zero + IPC handle, 64 bytes), pem_peer_open (a peer's handle -> device pointer valid in this process),
 *   pem_peer_release(ptr, opened).  peer_tables: HOST array of `world` device pointers, peer_tables[rank] =
 *   the own table.
 * A peer that does not deliver within timeout_ms sets word 1 of the own table's header instead of hanging.
 */
size_t pem_peer_table_bytes(int world, int n_per_rank);
int pem_peer_alloc(size_t bytes, void **ptr, unsigned char *ipc_handle);
int pem_peer_open(const unsigned char *ipc_handle, void **ptr);
int pem_peer_release(void *ptr, int opened);
int pem_gather_poses(const float *R, const float *t, const float *score, const int32_t *proposal_id,
                     const float *instance_id, int n_per_rank, int rank, int world, void *const *peer_tables,
                     float *records_local, pem_stream_t stream);
int pem_gather_poses_wait(void *own_table, int world, int n_per_rank, float *out, int timeout_ms,
                          pem_stream_t stream);

/* fp16 -> fp32 widening of an uploaded tensor (host-facing path with half-precision features); n % 8 == 0. */
int pem_half_to_float(const void *src, float *dst, size_t n, pem_stream_t stream);

/*
 * Measurement aid (no reference counterpart; SURVEY 8d asks for it): an FP32 issue-peak probe.  `ctas` CTAs of
 * 512 threads each run `iters` x 64 dependent-chain instructions per lane over 8 independent chains: mode 0 =
 * fused multiply-adds (2 flop each), mode 1 = unfused FMUL / FADD (the instruction mix bit-exactness with the
 * reference's CPU loops mandates for FPS / ball query), mode 2 = packed fma.rn.f32x2 (4 flop each).  sink: one float, never written in practice.
 */
int pem_bench_fp32_peak(int mode, int ctas, int iters, float *sink, pem_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PEM_B200_H */
