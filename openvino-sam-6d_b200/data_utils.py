"""Mirror of the host-side instance preparation of the reference's inference script (SURVEY.md
section 8f, rank 4): utils/data_utils.py (get_point_cloud_from_depth :92-110, get_bbox :125-160,
get_resize_rgb_choose :112-123) and the geometric lines of get_test_data
(run_inference_custom_pytorch.py:295-361), plus the result wire format (:459-476).

The depth image, the decoded masks and the camera go to the device once; every detection is one CTA
(instance_prep.cu).  The only step left on the host is the random draw: the reference samples with
numpy's global generator, and re-using that generator -- same calls, same order -- is what makes a
seeded run return the same points as the reference.  The RGB crop / cv2.resize feeds the ViT and is
not part of the geometric path; `bboxes` and `choose` are returned so the caller can do it as before.
"""
import json

import numpy as np
import torch

from . import _lib
from . import pointnet2_ext as _ext


def numpy_precision():
    """Which arithmetic the reference's lines get from the installed numpy: "f32" under value-based
    casting (numpy < 2, the reference's pinned 1.26), "f64" under NEP 50 (NumPy >= 2)."""
    return "f32" if (np.ones(1, np.float32) - np.float64(0.5)).dtype == np.float32 else "f64"


def _prec(precision):
    precision = precision or numpy_precision()
    if precision not in ("f32", "f64"):
        raise ValueError("precision must be 'f32', 'f64' or None (follow the installed numpy)")
    return precision, (0 if precision == "f32" else 1)


def _camera(K):
    K = np.asarray(K, np.float64).reshape(3, 3)
    return float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])


def depth_to_metres(raw_depth, depth_scale, device="cuda"):
    """`load_im(depth_path).astype(np.float32) * cam_info['depth_scale'] / 1000.0` (:295) for a uint16
    sensor image: (H,W) uint16 (numpy or tensor) -> (H,W) float32 CUDA tensor in metres."""
    if isinstance(raw_depth, np.ndarray):
        # check BEFORE the reinterpreting view: an int32 / float32 / uint8 image (imageio and cv2 can return
        # those) must raise, not be read as a 16-bit image of another width
        if raw_depth.dtype.kind not in "iu" or raw_depth.dtype.itemsize != 2:
            raise TypeError("raw depth must be a 16-bit integer image, got %s" % raw_depth.dtype)
        raw = torch.as_tensor(np.ascontiguousarray(raw_depth).view(np.int16))
    else:
        raw = torch.as_tensor(raw_depth)
    if raw.dtype not in (torch.int16, torch.uint16):
        raise TypeError("raw depth must be a 16-bit integer image, got %s" % raw.dtype)
    if raw.ndim != 2:
        raise ValueError("raw depth must be (H,W)")
    raw = raw.to(device).contiguous()
    H, W = raw.shape
    out = torch.empty((H, W), dtype=torch.float32, device=raw.device)
    with torch.cuda.device(raw.device):
        _lib.check(_lib.lib().pem_depth_to_metres(raw.data_ptr(), H, W, float(depth_scale), out.data_ptr(),
                                                  _ext._stream()), "pem_depth_to_metres")
    return out


def get_point_cloud_from_depth(depth, K, bbox=None, precision=None):
    """depth (H,W) float32 CUDA tensor in metres, K (3,3) -> (H,W,3) float32 cloud (cropped to
    bbox = [rmin, rmax, cmin, cmax] when given, as the reference's optional argument does)."""
    _ext._check(depth, "depth", torch.float32, "a float")
    _, pflag = _prec(precision)
    fx, fy, cx, cy = _camera(K)
    H, W = depth.shape
    cloud = torch.empty((H, W, 3), dtype=torch.float32, device=depth.device)
    with torch.cuda.device(depth.device):
        _lib.check(_lib.lib().pem_backproject_depth(depth.data_ptr(), H, W, fx, fy, cx, cy, pflag,
                                                    cloud.data_ptr(), _ext._stream()), "pem_backproject_depth")
    if bbox is not None:
        rmin, rmax, cmin, cmax = bbox
        cloud = cloud[rmin:rmax, cmin:cmax]
    return cloud


def radius_threshold(radius, precision=None):
    """`radius * 1.2` with radius = np.max(np.linalg.norm(model_points, axis=1)) a float32 scalar (:300,
    :327), as the comparison sees it: numpy < 2 multiplies in float64 and compares in float32 after
    rounding the scalar, NumPy >= 2 multiplies in float32."""
    precision, _ = _prec(precision)
    r = np.float32(radius)
    return float(np.float32(float(r) * 1.2)) if precision == "f32" else float(np.float32(r * np.float32(1.2)))


def prepare_instances(depth, K, masks, scores, radius, n_sample_observed_point=2048, img_size=224,
                      rng=None, precision=None):
    """The detection loop of get_test_data (:303-361) minus the RGB crop.

    depth (H,W) float32 CUDA metres; masks (D,H,W) uint8/bool CUDA (decoded, row-major); scores (D,);
    radius: the model's radius (np.max(np.linalg.norm(model_points, axis=1))).
    Returns a dict with, for the n kept detections in order: 'pts' (n,ns,3) f32, 'rgb_choose' (n,ns) i64,
    'choose' (n,ns) i64 crop-flat pixel indices, 'score' (n,) f32 (all CUDA), 'kept' (list of detection
    numbers), 'bboxes' (n,4) [y1,y2,x1,x2] and 'n_candidates' (n,) numpy.  Detections with <= 32 masked
    pixels or < 4 points within 1.2 radii of their mean are dropped, as in the reference."""
    _ext._check(depth, "depth", torch.float32, "a float")
    precision, pflag = _prec(precision)
    rng = np.random if rng is None else rng
    if masks.dtype == torch.bool:
        masks = masks.to(torch.uint8)
    if masks.dtype != torch.uint8 or not masks.is_cuda or masks.dim() != 3:
        raise TypeError("masks must be a (D,H,W) uint8 CUDA tensor")
    masks = masks.contiguous()
    D, H, W = masks.shape
    if tuple(depth.shape) != (H, W):
        raise ValueError("depth and masks disagree on the image size")
    fx, fy, cx, cy = _camera(K)
    L, dev = _lib.lib(), depth.device
    nbytes = L.pem_instance_scratch_bytes(D, H, W)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    info = torch.empty((D, 8), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        st = _ext._stream()
        _lib.check(L.pem_instance_candidates(depth.data_ptr(), masks.data_ptr(), D, H, W, fx, fy, cx, cy, pflag,
                                             radius_threshold(radius, precision), info.data_ptr(),
                                             scratch.data_ptr(), nbytes, st), "pem_instance_candidates")
        h_info = info.cpu().numpy()  # 32 bytes per detection: the host needs the counts to draw
        kept = [d for d in range(D) if h_info[d, 0] == 1]
        if not kept:
            raise RuntimeError("no detection survived the preparation (the reference fails in torch.stack here)")
        ns = int(n_sample_observed_point)
        picks = []
        for d in kept:  # np.random.choice, with replacement when there are not enough points (:333-336)
            m = int(h_info[d, 2])
            picks.append(rng.choice(np.arange(m), ns) if m <= ns else rng.choice(np.arange(m), ns, replace=False))
        pick = torch.from_numpy(np.stack(picks).astype(np.int64)).to(dev)
        det_ids = torch.tensor(kept, dtype=torch.int32, device=dev)
        n = len(kept)
        pts = torch.empty((n, ns, 3), dtype=torch.float32, device=dev)
        rgb_choose = torch.empty((n, ns), dtype=torch.int64, device=dev)
        choose = torch.empty((n, ns), dtype=torch.int64, device=dev)
        _lib.check(L.pem_instance_sample(depth.data_ptr(), H, W, fx, fy, cx, cy, pflag, info.data_ptr(),
                                         scratch.data_ptr(), nbytes, D, det_ids.data_ptr(), n, pick.data_ptr(),
                                         ns, int(img_size), pts.data_ptr(), rgb_choose.data_ptr(),
                                         choose.data_ptr(), st), "pem_instance_sample")
    score = torch.as_tensor(np.asarray(scores, np.float64)[kept], dtype=torch.float32).to(dev)
    return {"pts": pts, "rgb_choose": rgb_choose, "choose": choose, "score": score, "kept": kept,
            "bboxes": h_info[kept, 3:7].copy(), "n_candidates": h_info[kept, 2].copy()}


def detections_to_json(detections, pred_R, pred_t, pose_scores):
    """The result wire format, detection_pem_*.json (:467-476): every kept detection dict of the ISM
    gets 'score' (float), 'R' (3x3 nested list) and 't' (list, MILLIMETRES = pred_t * 1000) and the list
    is dumped with json.dump's defaults.  Which score goes in is the caller's choice, as in the
    reference: the torch script writes the ISM score (its `'pred_pose_score' in input_data` test at
    :460-463 is never true), the OpenVINO scripts write pose_score * ISM score
    (run_inference_custom_openvino_gpu.py:636)."""
    R = np.asarray(pred_R.detach().cpu() if torch.is_tensor(pred_R) else pred_R)
    t = np.asarray(pred_t.detach().cpu() if torch.is_tensor(pred_t) else pred_t) * 1000
    s = np.asarray(pose_scores.detach().cpu() if torch.is_tensor(pose_scores) else pose_scores)
    out = []
    for i, det in enumerate(detections):
        det = dict(det)
        det["score"] = float(s[i])
        det["R"] = list(R[i].tolist())
        det["t"] = list(t[i].tolist())
        out.append(det)
    return json.dumps(out)
