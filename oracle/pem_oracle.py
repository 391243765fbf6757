"""oracle/pem_oracle.py -- CPU restatement of the PEM geometric hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file;
it is used by tests/, by ``__graft_entry__.smoke()`` as the checker, and by the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

Two halves:

* the pointnet2 ops are the plain-C restatement in ``pointnet2_oracle.c``
  (built by ``oracle/Makefile`` into ``oracle/_build/libpem_oracle.so``) called
  through ctypes on numpy arrays;
* the dense geometric functions are restated with torch *CPU* ops, because the
  reference itself is torch CPU code whose third-party arithmetic (ATen
  softmax/cumsum, LAPACK gesdd through ``torch.svd``, ``torch.det``,
  ``torch.topk``) cannot be re-derived bit-for-bit in numpy.  The restatement
  is staged (every intermediate the parity plan S1..S5 needs is returned) and
  takes the random draw as an explicit argument.

Pinning: validated against the live, unmodified reference
(``oracle/ref_live.py``, this container only) by ``tests/test_oracle_cpu.py``
and against the committed fixtures in ``tests/golden/`` written by
``oracle/make_golden.py`` from that live reference.

Reference citations are relative to
/root/reference/SAM-6D/Pose_Estimation_Model/ (``MU`` = utils/model_utils.py).
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


# --------------------------------------------------------------------------
# pointnet2 ops: C oracle through ctypes
# --------------------------------------------------------------------------
def lib_path():
    return os.path.join(_HERE, "_build", "libpem_oracle.so")


def build_c_oracle():
    """gcc the C restatement (idempotent)."""
    src = os.path.join(_HERE, "pointnet2_oracle.c")
    out = lib_path()
    if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(
        ["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, src])
    return out


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c_oracle())
    return _LIB


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def fps(xyz, npoint, skip_near_origin=True):
    """(B,N,3) f32 -> (B,npoint) i32.  sampling.cpp:76-118 (CPU semantics)."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    idx = np.zeros((B, npoint), np.int32)
    scratch = np.empty((B, N), np.float32)
    _lib().oracle_fps(_p(xyz), B, N, npoint, int(skip_near_origin), _p(scratch), _p(idx))
    return idx


def gather(feat, idx):
    """(B,C,N) f32, (B,M) i32 -> (B,C,M).  sampling.cpp:23-44."""
    feat, idx = _f32(feat), _i32(idx)
    B, C, N = feat.shape
    M = idx.shape[1]
    out = np.zeros((B, C, M), np.float32)
    _lib().oracle_gather(_p(feat), _p(idx), B, C, N, M, _p(out))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """(B,M,3),(B,N,3) -> (B,M,nsample) i32.  ball_query.cpp:16-62."""
    new_xyz, xyz = _f32(new_xyz), _f32(xyz)
    B, M, _ = new_xyz.shape
    N = xyz.shape[1]
    idx = np.zeros((B, M, nsample), np.int32)
    _lib().oracle_ball_query(_p(new_xyz), _p(xyz), B, M, N,
                             ctypes.c_float(radius), nsample, _p(idx))
    return idx


def group(feat, idx):
    """(B,C,N) f32, (B,M,ns) i32 -> (B,C,M,ns).  group_points.cpp:20-45."""
    feat, idx = _f32(feat), _i32(idx)
    B, C, N = feat.shape
    _, M, ns = idx.shape
    out = np.zeros((B, C, M, ns), np.float32)
    _lib().oracle_group(_p(feat), _p(idx), B, C, N, M, ns, _p(out))
    return out


def query_and_group(xyz, new_xyz, features, radius, nsample):
    """QueryAndGroup.forward, use_xyz=True (pointnet2_utils.py:326-403).

    xyz (B,N,3), new_xyz (B,M,3), features (B,C,N) -> (B,3+C,M,ns), idx.
    """
    xyz, new_xyz, features = _f32(xyz), _f32(new_xyz), _f32(features)
    idx = ball_query(new_xyz, xyz, radius, nsample)
    g_xyz = group(np.ascontiguousarray(xyz.transpose(0, 2, 1)), idx)
    g_xyz = g_xyz - new_xyz.transpose(0, 2, 1)[:, :, :, None]
    g_feat = group(features, idx)
    return np.concatenate([g_xyz, g_feat], axis=1), idx


def sample_pts_feats(pts, feats, npoint):
    """MU:70-84.  pts (B,N,3), feats (B,N,C) -> (B,npoint,3), (B,npoint,C), idx."""
    pts, feats = _f32(pts), _f32(feats)
    idx = fps(pts, npoint)
    p = gather(np.ascontiguousarray(pts.transpose(0, 2, 1)), idx).transpose(0, 2, 1)
    f = gather(np.ascontiguousarray(feats.transpose(0, 2, 1)), idx).transpose(0, 2, 1)
    return p, f, idx


# --------------------------------------------------------------------------
# dense geometric functions: torch CPU restatement
# --------------------------------------------------------------------------
def _t(a):
    return a if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a))


def feature_similarity(feat1, feat2, temp=1.0, normalize=True):
    """Cosine similarity / temp.  MU:131-153 ('cosine' branch; F.normalize eps=1e-12)."""
    a, b = _t(feat1).float(), _t(feat2).float()
    if normalize:
        a = a / a.norm(dim=2, keepdim=True).clamp_min(1e-12)
        b = b / b.norm(dim=2, keepdim=True).clamp_min(1e-12)
    return torch.bmm(a, b.transpose(1, 2)) / temp


def sq_pairwise(x, y):
    """x2 - 2xy + y2 clamped at 0 (MU:101-128, expanded form, not sum((x-y)^2))."""
    xy = x @ y.transpose(-1, -2)
    x2 = (x ** 2).sum(-1).unsqueeze(-1)
    y2 = (y ** 2).sum(-1).unsqueeze(-2)
    return (x2 - 2 * xy + y2).clamp(min=0.0)


def dual_softmax(atten):
    """softmax over columns times softmax over rows (MU:230 / MU:320)."""
    atten = _t(atten).float()
    return torch.softmax(atten, dim=2) * torch.softmax(atten, dim=1)


def fg_masks(P):
    """label>0 masks (MU:231-234 / MU:321-322): argmax over the full row (col) for
    rows (cols) 1.., foreground iff the first maximum is not at index 0."""
    lab1 = P[:, 1:, :].max(dim=2)[1]
    lab2 = P[:, :, 1:].max(dim=1)[1]
    return (lab1 > 0).float(), (lab2 > 0).float()


def coarse_assignment(atten):
    """MU:230-242.  Returns w1 (B,N1), w2 (B,N2) and the normalised CDF (B,N1*N2) f32."""
    P = dual_softmax(atten)
    w1, w2 = fg_masks(P)
    B = P.shape[0]
    S = P[:, 1:, 1:].contiguous() * w1.unsqueeze(2) * w2.unsqueeze(1)
    S = S.reshape(B, -1) ** 1.5
    cdf = torch.cumsum(S, dim=1)
    cdf = cdf / (cdf[:, -1].unsqueeze(1).contiguous() + 1e-8)
    return {"w1": w1, "w2": w2, "score": S, "cdf": cdf}


def inverse_cdf(cdf, rand, faithful=False):
    """First index i with cdf[i] >= r, and 0 when there is none (MU:277-305).

    faithful=True materialises the (B,S,N) comparison exactly like the reference
    (used for the CPU baseline timing); the default is the equivalent
    lower-bound search (cdf is non-decreasing), patched for the "none" case.
    """
    cdf, rand = _t(cdf), _t(rand)
    if faithful:
        return torch.argmax((cdf.unsqueeze(1) >= rand.unsqueeze(2)).float(), dim=2)
    idx = torch.searchsorted(cdf.contiguous(), rand.contiguous(), right=False)
    return torch.where(idx >= cdf.shape[1], torch.zeros_like(idx), idx)


def procrustes(src, ref, weights=None, weight_thresh=0.0, eps=1e-5):
    """Weighted Kabsch, ref ~ R src + t.  MU:343-436."""
    src, ref = _t(src).float(), _t(ref).float()
    n = src.shape[0]
    w = torch.ones_like(src[:, :, 0]) if weights is None else _t(weights).float()
    w = torch.where(w < weight_thresh, torch.zeros_like(w), w)
    w = (w / (w.sum(dim=1, keepdim=True) + eps)).unsqueeze(2)
    cs = (src * w).sum(dim=1, keepdim=True)
    cr = (ref * w).sum(dim=1, keepdim=True)
    H = (src - cs).permute(0, 2, 1) @ (w * (ref - cr))
    U, _, V = torch.svd(H)
    Ut = U.transpose(1, 2)
    D = torch.eye(3).unsqueeze(0).repeat(n, 1, 1)
    D[:, 2, 2] = torch.sign(torch.det(V @ Ut))
    R = V @ D @ Ut
    t = (cr.permute(0, 2, 1) - R @ cs.permute(0, 2, 1)).squeeze(2)
    return R, t, H


def coarse_hypotheses(cdf, rand, pts1, pts2, n_proposal1, faithful=False):
    """MU:248-261: draw 3*P1 correspondences, solve P1 three-point Procrustes, residual."""
    pts1, pts2 = _t(pts1).float(), _t(pts2).float()
    B, N1, _ = pts1.shape
    N2 = pts2.shape[1]
    idx = inverse_cdf(cdf, rand, faithful)
    i1 = torch.clamp(idx.div(N2, rounding_mode="floor"), max=N1 - 1)
    i2 = torch.clamp(idx % N2, max=N2 - 1)
    p1 = torch.gather(pts1, 1, i1.unsqueeze(2).repeat(1, 1, 3)).reshape(B * n_proposal1, 3, 3)
    p2 = torch.gather(pts2, 1, i2.unsqueeze(2).repeat(1, 1, 3)).reshape(B * n_proposal1, 3, 3)
    R, t, H = procrustes(p2, p1, None, weight_thresh=0.5)
    R = R.reshape(B, n_proposal1, 3, 3)
    t = t.reshape(B, n_proposal1, 1, 3)
    p1 = p1.reshape(B, n_proposal1, 3, 3)
    p2 = p2.reshape(B, n_proposal1, 3, 3)
    resid = torch.norm((p1 - t) @ R - p2, dim=3).mean(2)
    return {"idx": idx, "i1": i1, "i2": i2, "R": R, "t": t.squeeze(2), "resid": resid,
            "H": H.reshape(B, n_proposal1, 3, 3)}


def score_hypotheses(R, t, pts1, w1, model_pts):
    """MU:267-271: score = sum(w1) / (sum_i w1_i * min_m ||(p1_i - t) R - m|| + 1e-8).

    R (B,H,3,3), t (B,H,3), pts1 (B,N1,3), w1 (B,N1), model_pts (B,Nm,3) -> (B,H).
    Chunked over H so large sweeps fit in memory.
    """
    R, t = _t(R).float(), _t(t).float()
    pts1, w1, model_pts = _t(pts1).float(), _t(w1).float(), _t(model_pts).float()
    B, Hn = R.shape[:2]
    out = torch.empty(B, Hn)
    step = 256
    for h0 in range(0, Hn, step):
        Rc, tc = R[:, h0:h0 + step], t[:, h0:h0 + step]
        tp = (pts1.unsqueeze(1) - tc.unsqueeze(2)) @ Rc            # (B,h,N1,3)
        h = tp.shape[1]
        tp = tp.reshape(B * h, -1, 3)
        mp = model_pts.unsqueeze(1).repeat(1, h, 1, 1).reshape(B * h, -1, 3)
        dis = torch.sqrt(sq_pairwise(tp, mp)).min(2)[0].reshape(B, h, -1)
        out[:, h0:h0 + step] = w1.unsqueeze(1).sum(2) / ((dis * w1.unsqueeze(1)).sum(2) + 1e-8)
    return out


def coarse_Rt(atten, pts1, pts2, model_pts, n_proposal1, n_proposal2, rand, faithful=False):
    """compute_coarse_Rt (MU:204-275) with the random draw made explicit.

    Returns (R (B,3,3), t (B,3), stages) where ``stages`` holds every intermediate
    of the parity plan plus ``proposal_id`` = original index (in [0,P1)) of the
    selected hypothesis.
    """
    pts1, pts2 = _t(pts1).float(), _t(pts2).float()
    model_pts = pts2 if model_pts is None else _t(model_pts).float()
    B = pts1.shape[0]
    a = coarse_assignment(atten)
    h = coarse_hypotheses(a["cdf"], rand, pts1, pts2, n_proposal1, faithful)
    top = torch.topk(h["resid"], n_proposal2, dim=1, largest=False)[1]
    Rs = torch.gather(h["R"], 1, top.reshape(B, -1, 1, 1).repeat(1, 1, 3, 3))
    ts = torch.gather(h["t"], 1, top.reshape(B, -1, 1).repeat(1, 1, 3))
    scores = score_hypotheses(Rs, ts, pts1, a["w1"], model_pts)
    best = scores.max(1)[1]
    ar = torch.arange(B)
    stages = dict(a)
    stages.update(h)
    stages.update({"top": top, "scores": scores, "best": best, "proposal_id": top[ar, best]})
    return Rs[ar, best], ts[ar, best], stages


def fine_Rt(atten, pts1, pts2, model_pts=None, dis_thres=0.15):
    """compute_fine_Rt (MU:308-341).  Returns R, t, pose_score, stages."""
    pts1, pts2 = _t(pts1).float(), _t(pts2).float()
    model_pts = pts2 if model_pts is None else _t(model_pts).float()
    P = dual_softmax(atten)
    w1, w2 = fg_masks(P)
    A = P[:, 1:, 1:] * w1.unsqueeze(2) * w2.unsqueeze(1)
    row = A.sum(2)
    pred = (A / (A.sum(2, keepdim=True) + 1e-6)) @ pts2
    R, t, H = procrustes(pred, pts1, row, weight_thresh=0.0)
    moved = (pts1 - t.unsqueeze(1)) @ R
    dis = torch.sqrt(sq_pairwise(moved, model_pts)).min(2)[0]
    hit = (dis < dis_thres).float()
    score = (hit * w1).sum(1) / (w1.sum(1) + 1e-8)
    score = score * w1.mean(1)
    return R, t, score, {"w1": w1, "w2": w2, "row": row, "pred": pred, "H": H, "dis": dis}


# --------------------------------------------------------------------------
# comparison helpers shared by tests, smoke() and bench
# --------------------------------------------------------------------------
def geo_embedding_indices(points, sigma_d=0.2, sigma_a=15.0, angle_k=3):
    """GeometricStructureEmbedding.get_embedding_indices (model/transformer.py:304-341): indices of
    the pair-wise distance embedding and of the triplet-wise angular embedding.

    points (B,N,3) -> d_indices (B,N,N) = dist / sigma_d, a_indices (B,N,N,k) = angle * 180/(sigma_a*pi),
    angle between (knn_l(i) - p_i) and (p_j - p_i); the k nearest neighbours are entries 1..k of the
    (k+1) smallest distances of row i (:324), sin clamped at 1e-8 and the raw dot product clamped to
    (-1+1e-8, 1-1e-8), i.e. [-1, 1] in fp32 (:338-340).
    """
    points = _t(points).float()
    B, N, _ = points.shape
    dist = torch.sqrt(sq_pairwise(points, points))                                   # :321
    d_indices = dist / sigma_d                                                       # :322
    k = angle_k
    knn = dist.topk(k=k + 1, dim=2, largest=False)[1][:, :, 1:]                      # :325 (B,N,k)
    knn_points = torch.gather(points.unsqueeze(1).expand(B, N, N, 3), 2, knn.unsqueeze(3).expand(B, N, k, 3))
    ref = (knn_points - points.unsqueeze(2)).unsqueeze(2).expand(B, N, N, k, 3)      # :329,331
    anc = (points.unsqueeze(1) - points.unsqueeze(2)).unsqueeze(3).expand(B, N, N, k, 3)  # :330,332
    sin_v = torch.linalg.norm(torch.cross(ref, anc, dim=-1), dim=-1)                 # :333
    cos_v = torch.sum(ref * anc, dim=-1)                                             # :334
    eps = 1e-8
    sin_v = torch.clamp(sin_v, min=eps)                                              # :339
    cos_v = torch.clamp(cos_v, min=-1.0 + eps, max=1.0 - eps)                        # :340
    a_indices = torch.atan2(sin_v, cos_v) * (180.0 / (sigma_a * np.pi))              # :342-343
    return d_indices, a_indices, knn


def pe_shared_mlp_amax(grouped, layers):
    """PositionalEncoding's per-scale branch after the grouping (model/fine_point_matching.py:131-133,
    139-141): SharedMLP([6,32,64,128], bn=True) = three times [1x1 Conv2d without bias ->
    BatchNorm2d (eval) -> ReLU] (pointnet2/pytorch_utils.py:25-50,80-134), then amax over the
    neighbours.  grouped (B,6,N,ns); layers = [(conv_w (Co,Ci), bn_w, bn_b, run_mean, run_var, eps)] * 3
    -> (B,128,N)."""
    x = _t(grouped).float()
    for (w, g, b, m, v, eps) in layers:
        w, g, b, m, v = (_t(z).float() for z in (w, g, b, m, v))
        x = torch.nn.functional.conv2d(x, w.reshape(w.shape[0], w.shape[1], 1, 1))
        x = torch.nn.functional.batch_norm(x, m, v, g, b, False, 0.0, float(eps))
        x = torch.relu(x)
    return torch.amax(x, dim=3)


def positional_encoding(pts, branch1, branch2, w3, b3, r1=0.1, ns1=32, r2=0.2, ns2=64):
    """PositionalEncoding.forward (model/fine_point_matching.py:113-144) on one cloud: both
    QueryAndGroup scales with centre = pts + 1e-8, the two SharedMLP + amax branches, concatenation
    and the final Conv1d(256, out_dim) without activation.  pts (B,N,3) -> (B,N,out_dim)."""
    pts = _t(pts).float().contiguous()
    ctr = (pts + 0.00000001).contiguous()
    feat = pts.transpose(1, 2).contiguous()
    g1, _ = query_and_group(pts.numpy(), ctr.numpy(), feat.numpy(), r1, ns1)
    g2, _ = query_and_group(pts.numpy(), ctr.numpy(), feat.numpy(), r2, ns2)
    f1 = pe_shared_mlp_amax(torch.from_numpy(g1), branch1)
    f2 = pe_shared_mlp_amax(torch.from_numpy(g2), branch2)
    x = torch.cat([f1, f2], dim=1)                                                   # (B,256,N)
    y = torch.nn.functional.conv1d(x, _t(w3).float().reshape(w3.shape[0], w3.shape[1], 1), _t(b3).float())
    return y.transpose(1, 2), f1, f2


def rotation_angle_deg(Ra, Rb):
    """Geodesic angle via the chord formula 2*asin(||Ra-Rb||_F / (2*sqrt(2))) in fp64
    (fp32 acos((tr-1)/2) bottoms out near 0.03 deg)."""
    Ra = _t(Ra).double()
    Rb = _t(Rb).double()
    chord = (Ra - Rb).flatten(-2).norm(dim=-1) / (2.0 * np.sqrt(2.0))
    return torch.rad2deg(2.0 * torch.asin(chord.clamp(max=1.0)))


# ---- host-side instance preparation -------------------------------
def _prep_dtype(precision):
    """The reference pins numpy 1.26 (SAM-6D/requirements.txt:32), where a float32 array combined with
    the float64 camera scalars stays float32 ("f32"); under NumPy >= 2 (NEP 50, this container) the
    same lines promote to float64 ("f64").  Both are restated; K's dtype selects the arithmetic."""
    assert precision in ("f32", "f64")
    return np.float32 if precision == "f32" else np.float64


def point_cloud_from_depth(depth, K, precision="f64"):
    """get_point_cloud_from_depth (utils/data_utils.py:92-110) without a bbox: (H,W) metres -> (H,W,3).
    pt0 = (x - cx) * z / fx, pt1 = (y - cy) * z / fy, pt2 = z, one rounding per operation."""
    dt = _prep_dtype(precision)
    K = np.asarray(K).astype(dt)
    H, W = depth.shape
    xmap = np.broadcast_to(np.arange(W)[None, :], (H, W)).astype(np.float32)
    ymap = np.broadcast_to(np.arange(H)[:, None], (H, W)).astype(np.float32)
    z = depth.astype(np.float32)
    pt0 = (xmap - K[0, 2]) * z / K[0, 0]
    pt1 = (ymap - K[1, 2]) * z / K[1, 1]
    return np.stack([pt0, pt1, z.astype(pt0.dtype)]).transpose((1, 2, 0))


def mask_bbox(label):
    """get_bbox (utils/data_utils.py:125-160): tight box of the mask, made square with side
    min(max(rows, cols), min(H, W)) around the integer centre, shifted back inside the image."""
    H, W = label.shape
    rows = np.where(np.any(label, axis=1))[0]
    cols = np.where(np.any(label, axis=0))[0]
    rmin, rmax, cmin, cmax = int(rows[0]), int(rows[-1]) + 1, int(cols[0]), int(cols[-1]) + 1
    b = min(max(rmax - rmin, cmax - cmin), min(H, W))
    cr, cc = int((rmin + rmax) / 2), int((cmin + cmax) / 2)
    half = int(b / 2)
    rmin, rmax, cmin, cmax = cr - half, cr + half, cc - half, cc + half
    if rmin < 0:
        rmax, rmin = rmax - rmin, 0
    if cmin < 0:
        cmax, cmin = cmax - cmin, 0
    if rmax > H:
        rmin, rmax = rmin - (rmax - H), H
    if cmax > W:
        cmin, cmax = cmin - (cmax - W), W
    return [rmin, rmax, cmin, cmax]


def resize_rgb_choose(choose, bbox, img_size):
    """get_resize_rgb_choose (utils/data_utils.py:112-123): crop-flat pixel index -> index in the
    img_size x img_size resized crop, floor(row * img_size / crop_h) * img_size + floor(col * ...)."""
    rmin, rmax, cmin, cmax = bbox
    ratio_h, ratio_w = img_size / (rmax - rmin), img_size / (cmax - cmin)
    row, col = choose // (cmax - cmin), choose % (cmax - cmin)
    return (np.floor(row * ratio_h) * img_size + np.floor(col * ratio_w)).astype(np.int64)


def instance_candidates(depth, cloud, mask, dist_thresh):
    """The geometric part of get_test_data's per-detection loop up to the random draw
    (run_inference_custom_pytorch.py:311-331): mask & depth > 0, more than 32 pixels, square bbox,
    crop-flat indices of the masked pixels in row-major order, their points, the SEQUENTIAL mean
    numpy takes over axis 0, and the distance filter against dist_thresh = radius * 1.2; fewer than 4
    survivors drop the detection.  Returns None or (bbox, choose, cloud) of the survivors."""
    m = np.logical_and(mask > 0, depth > 0)
    if np.sum(m) <= 32:
        return None
    y1, y2, x1, x2 = bbox = mask_bbox(m)
    choose = m[y1:y2, x1:x2].astype(np.float32).flatten().nonzero()[0]
    pts = cloud[y1:y2, x1:x2, :].reshape(-1, 3)[choose, :]
    acc = np.zeros(3, pts.dtype)
    for i in range(pts.shape[0]):  # np.mean(axis=0) of a C-contiguous (n,3) array adds row after row
        acc = acc + pts[i]
    centre = acc / pts.dtype.type(pts.shape[0])
    d = pts - centre[None, :]
    s = d * d
    flag = np.sqrt((s[:, 0] + s[:, 1]) + s[:, 2]) < dist_thresh
    if np.sum(flag) < 4:
        return None
    return bbox, choose[flag], pts[flag]


def prepare_instances(depth, K, masks, scores, radius, n_sample, img_size, rng, precision="f64"):
    """get_test_data's detection loop (run_inference_custom_pytorch.py:303-361) minus the RGB crop:
    returns pts (n,n_sample,3) f32, rgb_choose (n,n_sample) i64, score (n,), kept detection indices,
    and the whole cloud.  `rng` is numpy's global-style generator (np.random or a RandomState); the
    draws happen in the reference's order so the same seed gives the same sample."""
    dt = _prep_dtype(precision)
    cloud = point_cloud_from_depth(depth, K, precision)
    radius = np.float32(radius)
    # `flag = norm < radius * 1.2`: numpy 1.26 multiplies in float64 and compares in float32 after
    # rounding the scalar; NumPy >= 2 multiplies in float32 and compares exactly
    thresh = np.float32(float(radius) * 1.2) if precision == "f32" else radius * np.float32(1.2)
    pts_all, rc_all, sc_all, kept = [], [], [], []
    for i in range(len(masks)):
        cand = instance_candidates(depth, cloud, masks[i], dt(thresh))
        if cand is None:
            continue
        bbox, choose, pts = cand
        if len(choose) <= n_sample:
            pick = rng.choice(np.arange(len(choose)), n_sample)
        else:
            pick = rng.choice(np.arange(len(choose)), n_sample, replace=False)
        pts_all.append(pts[pick].astype(np.float32))
        rc_all.append(resize_rgb_choose(choose[pick], bbox, img_size))
        sc_all.append(scores[i])
        kept.append(i)
    return (np.stack(pts_all), np.stack(rc_all), np.asarray(sc_all, np.float32), kept,
            cloud.reshape(-1, 3))
