"""Host-side mirror of the geometric functions of the reference's ``utils/model_utils.py``
(MU): same names, argument meaning and return values, backed by libpem_b200 kernels.

  sample_pts_feats(pts, feats, npoint, return_index)              MU:70-84
  compute_feature_similarity(feat1, feat2, type, temp, normalize) MU:131-153
  compute_coarse_Rt(atten, pts1, pts2, model_pts, n1, n2)         MU:204-275
  compute_fine_Rt(atten, pts1, pts2, model_pts, dis_thres)        MU:308-341
  weighted_procrustes / WeightedProcrustes                        MU:343-455

Differences a caller can see, all deliberate:
  * the random draw of compute_coarse_Rt (``torch.rand`` inside MU:293) can be passed in
    (``rand=``) so that two implementations consume identical values; when omitted it is
    drawn with ``torch.rand`` on the tensors' device, as the reference does;
  * ``return_stages=True`` additionally returns the intermediates the parity plan checks;
  * CUDA tensors only.
"""
import torch
import torch.nn as nn

from . import _lib
from . import pointnet2_ext as _ext
from .pointnet2_utils import furthest_point_sample, gather_rows


def _f32c(t, name):
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be a float tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (libpem_b200 has no CPU path)" % name)
    return t.contiguous()


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)


def sample_pts_feats(pts, feats, npoint=2048, return_index=False):
    """FPS + point/feature gather.  pts (B,N,3), feats (B,N,C) -> (B,npoint,3), (B,npoint,C)
    [, idx (B,npoint) i32].  The gathers read the point-major tensors directly."""
    pts = _f32c(pts, "pts")
    if feats.dtype != torch.float32 or not feats.is_cuda or feats.device != pts.device:
        raise RuntimeError("feats must be a float CUDA tensor on the device of pts")
    B, N, _ = pts.shape
    M = int(npoint)
    if B == 0 or M == 0:
        idx = furthest_point_sample(pts, npoint)
        out_p = gather_rows(pts, idx)
    else:  # the sampling kernel stores the points it chooses (pem_fps_gather): one launch fewer
        L = _lib.lib()
        idx = torch.empty((B, M), dtype=torch.int32, device=pts.device)
        out_p = torch.empty((B, M, 3), dtype=torch.float32, device=pts.device)
        nb = L.pem_fps_workspace_bytes(B, N)
        ws = _ws(nb, pts.device) if nb else None
        with torch.cuda.device(pts.device):
            _lib.check(L.pem_fps_gather(pts.data_ptr(), B, N, M, 1, idx.data_ptr(), out_p.data_ptr(),
                                        None if ws is None else ws.data_ptr(), nb, _ext._stream()),
                       "pem_fps_gather")
    out_f = gather_rows(feats, idx)
    if return_index:
        return out_p, out_f, idx
    return out_p, out_f


def compute_feature_similarity(feat1, feat2, type="cosine", temp=1.0, normalize_feat=True,
                               backend="fp32"):
    """(B,N,C),(B,M,C) -> (B,N,M) cosine similarity / temp.

    backend="fp32": register-tiled FFMA kernel (exact fp32 arithmetic; used for the coarse
    stage).  backend="tc": split-bf16 tcgen05 GEMM (tensor cores, logits within ~1e-5)."""
    if type != "cosine":
        raise NotImplementedError("only the 'cosine' similarity of the PEM configs is provided")
    f1, f2 = _f32c(feat1, "feat1"), _f32c(feat2, "feat2")
    B, N1, C = f1.shape
    N2 = f2.shape[1]
    out = torch.empty((B, N1, N2), dtype=torch.float32, device=f1.device)
    L = _lib.lib()
    if backend == "tc":
        nbytes = L.pem_fine_fused_workspace_bytes(B, N1 - 1, N2 - 1, C)
        ws = _ws(nbytes, f1.device)
        with torch.cuda.device(f1.device):
            _lib.check(L.pem_feature_similarity_tc(f1.data_ptr(), f2.data_ptr(), B, N1, N2, C,
                                                   float(temp), 1 if normalize_feat else 0,
                                                   out.data_ptr(), ws.data_ptr(), nbytes,
                                                   _ext._stream()), "pem_feature_similarity_tc")
        return out
    nbytes = L.pem_similarity_workspace_bytes(B, N1, N2)
    ws = _ws(nbytes, f1.device)
    with torch.cuda.device(f1.device):
        _lib.check(L.pem_feature_similarity(f1.data_ptr(), f2.data_ptr(), B, N1, N2, C, float(temp),
                                            1 if normalize_feat else 0, out.data_ptr(),
                                            ws.data_ptr(), nbytes, _ext._stream()),
                   "pem_feature_similarity")
    return out


def compute_coarse_Rt(atten, pts1, pts2, model_pts=None, n_proposal1=6000, n_proposal2=300,
                      rand=None, return_stages=False):
    """atten (B,N1+1,N2+1), pts1 (B,N1,3), pts2 (B,N2,3), model_pts (B,Nm,3) -> R (B,3,3), t (B,3)."""
    atten, pts1, pts2 = _f32c(atten, "atten"), _f32c(pts1, "pts1"), _f32c(pts2, "pts2")
    model_pts = pts2 if model_pts is None else _f32c(model_pts, "model_pts")
    B, N1, _ = pts1.shape
    N2, Nm = pts2.shape[1], model_pts.shape[1]
    P1, P2 = int(n_proposal1), int(n_proposal2)
    dev = pts1.device
    if rand is None:
        rand = torch.rand(B, P1 * 3, device=dev)
    rand = _f32c(rand, "rand")
    f32 = dict(dtype=torch.float32, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    w1, w2 = torch.empty((B, N1), **f32), torch.empty((B, N2), **f32)
    cdf = torch.empty((B, N1 * N2), **f32)
    Rh, th = torch.empty((B, P1, 3, 3), **f32), torch.empty((B, P1, 3), **f32)
    resid = torch.empty((B, P1), **f32)
    corr = torch.empty((B, P1 * 3), **i32) if return_stages else None
    sel = torch.empty((B, P2), **i32)
    scores = torch.empty((B, P2), **f32)
    R, t = torch.empty((B, 3, 3), **f32), torch.empty((B, 3), **f32)
    pid = torch.empty((B,), **i32)
    L = _lib.lib()
    with torch.cuda.device(dev):
        st = _ext._stream()  # the current stream of the tensors' device (taken inside the device context)
        _lib.check(L.pem_coarse_assign(atten.data_ptr(), B, N1, N2, w1.data_ptr(), w2.data_ptr(),
                                       cdf.data_ptr(), st), "pem_coarse_assign")
        _lib.check(L.pem_coarse_hypotheses(cdf.data_ptr(), rand.data_ptr(), pts1.data_ptr(),
                                           pts2.data_ptr(), B, N1, N2, P1, Rh.data_ptr(),
                                           th.data_ptr(), resid.data_ptr(),
                                           None if corr is None else corr.data_ptr(), st),
                   "pem_coarse_hypotheses")
        _lib.check(L.pem_select_smallest(resid.data_ptr(), B, P1, P2, sel.data_ptr(), st),
                   "pem_select_smallest")
        _lib.check(L.pem_score_hypotheses(Rh.data_ptr(), th.data_ptr(), sel.data_ptr(), B, P1, P2,
                                          pts1.data_ptr(), w1.data_ptr(), N1, model_pts.data_ptr(),
                                          Nm, scores.data_ptr(), st), "pem_score_hypotheses")
        _lib.check(L.pem_select_best(scores.data_ptr(), sel.data_ptr(), Rh.data_ptr(),
                                     th.data_ptr(), B, P1, P2, R.data_ptr(), t.data_ptr(),
                                     pid.data_ptr(), st), "pem_select_best")
    if return_stages:
        return R, t, {"w1": w1, "w2": w2, "cdf": cdf, "corr": corr, "R": Rh, "t": th,
                      "resid": resid, "sel": sel, "scores": scores, "proposal_id": pid}
    return R, t


SCORE_KERNELS = {"auto": 0, "simt": 1, "tensor": 2}  # PEM_SCORE_* of include/pem_b200.h


def score_hypotheses(R, t, pts1, w1, model_pts, sel=None, kernel="auto"):
    """Standalone hypothesis scoring (MU:267-271) for the proposal sweep: R (B,P,3,3),
    t (B,P,3) -> scores (B,H) with H = P (sel None) or sel.shape[1].  kernel: "auto" (the tensor-core
    kernel for every shape it takes), "simt" or "tensor" (the parity tests compare the two)."""
    R, t = _f32c(R, "R"), _f32c(t, "t")
    pts1, w1, model_pts = _f32c(pts1, "pts1"), _f32c(w1, "w1"), _f32c(model_pts, "model_pts")
    B, P = R.shape[:2]
    H = P if sel is None else sel.shape[1]
    scores = torch.empty((B, H), dtype=torch.float32, device=R.device)
    with torch.cuda.device(R.device):
        _lib.check(_lib.lib().pem_score_hypotheses_ex(
            R.data_ptr(), t.data_ptr(), None if sel is None else sel.data_ptr(), B, P, H,
            pts1.data_ptr(), w1.data_ptr(), pts1.shape[1], model_pts.data_ptr(),
            model_pts.shape[1], scores.data_ptr(), SCORE_KERNELS[kernel], _ext._stream()), "pem_score_hypotheses")
    return scores


def compute_fine_Rt(atten, pts1, pts2, model_pts=None, dis_thres=0.15, return_stages=False):
    """atten (B,N1+1,N2+1), pts1 (B,N1,3), pts2 (B,N2,3) -> R (B,3,3), t (B,3), pose_score (B)."""
    atten, pts1, pts2 = _f32c(atten, "atten"), _f32c(pts1, "pts1"), _f32c(pts2, "pts2")
    model_pts = pts2 if model_pts is None else _f32c(model_pts, "model_pts")
    B, N1, _ = pts1.shape
    N2, Nm = pts2.shape[1], model_pts.shape[1]
    dev = pts1.device
    f32 = dict(dtype=torch.float32, device=dev)
    R, t, score = torch.empty((B, 3, 3), **f32), torch.empty((B, 3), **f32), torch.empty((B,), **f32)
    w1 = w2 = row = pred = None
    if return_stages:
        w1, w2 = torch.empty((B, N1), **f32), torch.empty((B, N2), **f32)
        row, pred = torch.empty((B, N1), **f32), torch.empty((B, N1, 3), **f32)
    L = _lib.lib()
    nbytes = L.pem_fine_workspace_bytes(B, N1, N2)
    ws = _ws(nbytes, dev)
    p = lambda x: None if x is None else x.data_ptr()
    with torch.cuda.device(dev):
        _lib.check(L.pem_fine_Rt(atten.data_ptr(), pts1.data_ptr(), pts2.data_ptr(),
                                 model_pts.data_ptr(), B, N1, N2, Nm, float(dis_thres),
                                 R.data_ptr(), t.data_ptr(), score.data_ptr(), p(w1), p(w2),
                                 p(row), p(pred), ws.data_ptr(), nbytes, _ext._stream()),
                   "pem_fine_Rt")
    if return_stages:
        return R, t, score, {"w1": w1, "w2": w2, "row": row, "pred": pred}
    return R, t, score


FINE_PRECISIONS = {"split_bf16": 0, "fp16": 1}  # PEM_PREC_* of include/pem_b200.h


def compute_fine_Rt_fused(feat1, feat2, pts1, pts2, model_pts=None, temp=0.1, dis_thres=0.15,
                          return_stages=False, precision="split_bf16", three_pass=False, steps_per_task=0):
    """compute_feature_similarity(feat1, feat2, 'cosine', temp, True) followed by
    compute_fine_Rt(...), fused: the attention matrix never exists in memory.
    feat1 (B,N1+1,C), feat2 (B,N2+1,C) carry the background token in row 0
    (fine_point_matching.py:48-51).  precision: "split_bf16" (three bf16 products per element,
    fp32-class logits) or "fp16" (one fp16 product, logits to ~1e-3 of +-1/temp).  three_pass /
    steps_per_task: PEM_FINE_THREE_PASS / PEM_FINE_STEPS_PER_TASK of include/pem_b200.h (same results by
    another schedule; the parity tests hold the forms against each other)."""
    if precision not in FINE_PRECISIONS:
        raise ValueError("precision must be one of %s" % sorted(FINE_PRECISIONS))
    f1, f2 = _f32c(feat1, "feat1"), _f32c(feat2, "feat2")
    pts1, pts2 = _f32c(pts1, "pts1"), _f32c(pts2, "pts2")
    model_pts = pts2 if model_pts is None else _f32c(model_pts, "model_pts")
    B, N1, _ = pts1.shape
    N2, Nm, C = pts2.shape[1], model_pts.shape[1], f1.shape[2]
    assert f1.shape[1] == N1 + 1 and f2.shape[1] == N2 + 1
    dev = pts1.device
    f32 = dict(dtype=torch.float32, device=dev)
    R, t, score = torch.empty((B, 3, 3), **f32), torch.empty((B, 3), **f32), torch.empty((B,), **f32)
    w1 = w2 = row = pred = None
    if return_stages:
        w1, w2 = torch.empty((B, N1), **f32), torch.empty((B, N2), **f32)
        row, pred = torch.empty((B, N1), **f32), torch.empty((B, N1, 3), **f32)
    L = _lib.lib()
    nbytes = L.pem_fine_fused_workspace_bytes(B, N1, N2, C)
    ws = _ws(nbytes, dev)
    p = lambda x: None if x is None else x.data_ptr()
    with torch.cuda.device(dev):
        _lib.check(L.pem_fine_fused_ex(f1.data_ptr(), f2.data_ptr(), pts1.data_ptr(), pts2.data_ptr(),
                                       model_pts.data_ptr(), B, N1, N2, C, Nm, float(temp),
                                       float(dis_thres), FINE_PRECISIONS[precision] | (0x100 if three_pass else 0) |
                                       ((int(steps_per_task) & 0xff) << 16), R.data_ptr(),
                                       t.data_ptr(), score.data_ptr(), p(w1), p(w2), p(row), p(pred),
                                       ws.data_ptr(), nbytes, _ext._stream()), "pem_fine_fused_ex")
    if return_stages:
        return R, t, score, {"w1": w1, "w2": w2, "row": row, "pred": pred}
    return R, t, score


def weighted_procrustes(src_points, ref_points, weights=None, weight_thresh=0.0, eps=1e-5,
                        return_transform=False):
    """MU:343-436: rigid transform with ref ~ R src + t.  (B,n,3) or (n,3) inputs."""
    squeeze = src_points.ndim == 2
    if squeeze:
        src_points, ref_points = src_points.unsqueeze(0), ref_points.unsqueeze(0)
        weights = None if weights is None else weights.unsqueeze(0)
    src, ref = _f32c(src_points, "src_points"), _f32c(ref_points, "ref_points")
    w = None if weights is None else _f32c(weights, "weights")
    B, n, _ = src.shape
    R = torch.empty((B, 3, 3), dtype=torch.float32, device=src.device)
    t = torch.empty((B, 3), dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        _lib.check(_lib.lib().pem_weighted_procrustes(
            src.data_ptr(), ref.data_ptr(), None if w is None else w.data_ptr(), B, n,
            float(weight_thresh), float(eps), R.data_ptr(), t.data_ptr(), _ext._stream()),
            "pem_weighted_procrustes")
    if return_transform:
        T = torch.eye(4, device=src.device).repeat(B, 1, 1)
        T[:, :3, :3], T[:, :3, 3] = R, t
        return T.squeeze(0) if squeeze else T
    return (R.squeeze(0), t.squeeze(0)) if squeeze else (R, t)


class WeightedProcrustes(nn.Module):
    """MU:438-455 (default weight_thresh 0.5, like the reference)."""

    def __init__(self, weight_thresh=0.5, eps=1e-5, return_transform=False):
        super().__init__()
        self.weight_thresh, self.eps, self.return_transform = weight_thresh, eps, return_transform

    def forward(self, src_points, tgt_points, weights=None):
        return weighted_procrustes(src_points, tgt_points, weights, self.weight_thresh, self.eps,
                                   self.return_transform)


# ---------------------------------------------------------------------------------------------
# The operator seam (MU:457-526): CustomSearchSorted / CustomSVD(u,v) / CustomDet.  Same class names,
# `apply` signatures, return values and ONNX `symbolic` names as the reference's shells; forward runs the
# batched libpem_b200 kernels instead of torch.searchsorted / torch.svd / torch.det.
# ---------------------------------------------------------------------------------------------
def searchsorted(cumsum_weights, random_values):
    """torch.searchsorted(cumsum_weights, random_values) (right=False) for (B,N) / (B,S) fp32 CUDA tensors
    -> (B,S) int64; also accepts 1-D inputs.  MU:462-463, custom_searchsorted.cpp:90-109."""
    one_d = cumsum_weights.ndim == 1
    seq = _f32c(cumsum_weights, "cumsum_weights")
    val = _f32c(random_values, "random_values")
    if one_d:
        seq, val = seq.unsqueeze(0), val.reshape(1, -1)
    if seq.ndim != 2 or val.ndim != 2 or seq.shape[0] != val.shape[0]:
        raise RuntimeError("searchsorted expects (B,N) and (B,S) tensors")
    B, N = seq.shape
    S = val.shape[1]
    out = torch.empty((B, S), dtype=torch.int64, device=seq.device)
    if out.numel():
        if N == 0:
            out.zero_()
        else:
            with torch.cuda.device(seq.device):
                _lib.check(_lib.lib().pem_searchsorted(seq.data_ptr(), val.data_ptr(), B, N, S, out.data_ptr(),
                                                       _ext._stream()), "pem_searchsorted")
    return out.reshape(random_values.shape) if one_d else out


def svd3(H, want=("U", "S", "V")):
    """Batched SVD of (...,3,3) fp32 CUDA matrices: H = U diag(S) V^T like torch.svd (MU:475)."""
    H = _f32c(H, "H")
    if H.shape[-2:] != (3, 3):
        raise RuntimeError("svd3 expects (...,3,3) matrices (the PEM pose solvers' shape)")
    n = H.numel() // 9
    mk = lambda *shape: torch.empty(H.shape[:-2] + shape, dtype=torch.float32, device=H.device)
    U = mk(3, 3) if "U" in want else None
    S = mk(3) if "S" in want else None
    V = mk(3, 3) if "V" in want else None
    if n:
        p = lambda x: None if x is None else x.data_ptr()
        with torch.cuda.device(H.device):
            _lib.check(_lib.lib().pem_svd3_batched(H.data_ptr(), n, p(U), p(S), p(V), _ext._stream()),
                       "pem_svd3_batched")
    return U, S, V


def det3(H):
    """Batched determinant of (...,3,3) fp32 CUDA matrices (MU:520)."""
    H = _f32c(H, "H")
    if H.shape[-2:] != (3, 3):
        raise RuntimeError("det3 expects (...,3,3) matrices")
    out = torch.empty(H.shape[:-2], dtype=torch.float32, device=H.device)
    if out.numel():
        with torch.cuda.device(H.device):
            _lib.check(_lib.lib().pem_det3_batched(H.data_ptr(), out.numel(), out.data_ptr(), _ext._stream()),
                       "pem_det3_batched")
    return out


class CustomSearchSorted(torch.autograd.Function):
    """MU:457-467."""

    @staticmethod
    def forward(ctx, cumsum_weights, random_values):
        return searchsorted(cumsum_weights, random_values)

    @staticmethod
    def symbolic(g, cumsum_weights, random_values):
        return g.op("CustomSearchSorted", cumsum_weights, random_values, outputs=1)


class CustomSVD(torch.autograd.Function):
    """MU:469-481: returns (U, S, V)."""

    @staticmethod
    def forward(ctx, H):
        U, S, V = svd3(H)
        ctx.save_for_backward(U, S, V)
        return U, S, V

    @staticmethod
    def symbolic(g, H):
        return g.op("CustomSVD", H, outputs=3)


class CustomSVDu(torch.autograd.Function):
    """MU:483-496."""

    @staticmethod
    def forward(ctx, H):
        return svd3(H, want=("U",))[0]

    @staticmethod
    def symbolic(g, H):
        return g.op("CustomSVDu", H, outputs=1)


class CustomSVDv(torch.autograd.Function):
    """MU:498-510."""

    @staticmethod
    def forward(ctx, H):
        return svd3(H, want=("V",))[2]

    @staticmethod
    def symbolic(g, H):
        return g.op("CustomSVDv", H, outputs=1)


class CustomDet(torch.autograd.Function):
    """MU:512-526."""

    @staticmethod
    def forward(ctx, H):
        det = det3(H)
        ctx.save_for_backward(det)
        return det

    @staticmethod
    def symbolic(g, H):
        return g.op("CustomDet", H, outputs=1)
