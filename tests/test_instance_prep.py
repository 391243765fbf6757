"""Host-side instance preparation (reference run_inference_custom_pytorch.py:295-361 and
utils/data_utils.py:92-160): the oracle against the golden vectors recorded
from the live get_test_data (oracle/make_golden_prep.py), and the CUDA path against the oracle and
the golden vectors -- everything bit exact (indices, and points that are one rounding per numpy op)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

G = os.path.join(GOLDEN, "instance_prep.npz")
NS, IMG = 2048, 224


def _golden():
    g = np.load(G)
    depth = g["depth_u16"].astype(np.float32) * 1.0 / 1000.0          # :295 with depth_scale = 1.0
    model = g["model_mm"].astype(np.float32) / 1000.0                  # :299
    radius = np.max(np.linalg.norm(model, axis=1))                     # :300
    return g, depth, g["K"].reshape(3, 3), radius


def test_oracle_matches_live_get_test_data(oracle):
    g, depth, K, radius = _golden()
    np.random.seed(11)
    pts, rc, sc, kept, whole = oracle.prepare_instances(depth, K, g["masks"], g["scores"], radius, NS, IMG,
                                                        np.random, "f64")
    assert kept == g["kept"].tolist()
    np.testing.assert_array_equal(pts, g["pts"])
    np.testing.assert_array_equal(rc, g["rgb_choose"])
    np.testing.assert_array_equal(sc, g["score"])
    np.testing.assert_array_equal(whole.astype(np.float32), g["whole_pts"])


def test_oracle_f32_backprojection_and_boxes(oracle):
    g, depth, K, _ = _golden()
    w32 = oracle.point_cloud_from_depth(depth, K, "f32")
    assert w32.dtype == np.float32
    np.testing.assert_array_equal(w32.reshape(-1, 3), g["whole_pts_f32"])
    assert (g["whole_pts_f32"] != g["whole_pts"]).mean() > 0.05           # the two arithmetics do differ
    m = np.logical_and(g["masks"] > 0, depth[None] > 0)
    assert [oracle.mask_bbox(m[i]) for i in range(len(m))] == g["bboxes"].tolist()


def test_detections_to_json(pkg):
    dets = [{"scene_id": 0, "image_id": 3, "category_id": 1, "bbox": [1, 2, 3, 4], "score": 0.5,
             "segmentation": {"size": [4, 4], "counts": "abc"}} for _ in range(2)]
    rs = np.random.RandomState(0)
    R, t, s = rs.rand(2, 3, 3).astype(np.float32), rs.rand(2, 3).astype(np.float32), rs.rand(2).astype(np.float32)
    ours = pkg.data_utils.detections_to_json(dets, torch.from_numpy(R), torch.from_numpy(t), torch.from_numpy(s))
    # the reference's statements (:467-476)
    ref = [dict(d) for d in dets]
    pred_trans = t * 1000
    for idx, det in enumerate(ref):
        ref[idx]["score"] = float(s[idx])
        ref[idx]["R"] = list(R[idx].tolist())
        ref[idx]["t"] = list(pred_trans[idx].tolist())
    assert ours == json.dumps(ref)
    assert dets[0]["score"] == 0.5 and "R" not in dets[0]                  # the caller's list is not modified


def test_precision_switch(pkg):
    du = pkg.data_utils
    assert du.numpy_precision() in ("f32", "f64")
    assert du.radius_threshold(0.17, "f32") == float(np.float32(float(np.float32(0.17)) * 1.2))
    assert du.radius_threshold(0.17, "f64") == float(np.float32(0.17) * np.float32(1.2))
    with pytest.raises(ValueError):
        du.radius_threshold(0.17, "f16")


# ---- GPU -----------------------------------------------------------------------------------
def _run(pkg, dev, depth, K, masks, scores, radius, precision, seed):
    d = torch.from_numpy(depth).to(dev)
    m = torch.from_numpy(np.ascontiguousarray(masks)).to(dev)
    rng = np.random.RandomState(seed)
    return pkg.data_utils.prepare_instances(d, K, m, scores, radius, NS, IMG, rng=rng, precision=precision)


@pytest.mark.gpu
def test_gpu_matches_golden_live_reference(pkg, dev):
    g, depth, K, radius = _golden()
    np.random.seed(11)
    d = torch.from_numpy(depth).to(dev)
    out = pkg.data_utils.prepare_instances(d, K, torch.from_numpy(g["masks"]).to(dev), g["scores"], radius, NS, IMG,
                                           rng=np.random, precision="f64")
    assert out["kept"] == g["kept"].tolist()
    np.testing.assert_array_equal(out["pts"].cpu().numpy(), g["pts"])
    np.testing.assert_array_equal(out["rgb_choose"].cpu().numpy(), g["rgb_choose"])
    np.testing.assert_array_equal(out["score"].cpu().numpy(), g["score"])
    np.testing.assert_array_equal(out["bboxes"], g["bboxes"][g["kept"]])


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_gpu_backprojection(pkg, dev, precision):
    g, depth, K, _ = _golden()
    raw = pkg.data_utils.depth_to_metres(g["depth_u16"], 1.0, dev)
    np.testing.assert_array_equal(raw.cpu().numpy(), depth)
    cloud = pkg.data_utils.get_point_cloud_from_depth(raw, K, precision=precision)
    want = g["whole_pts_f32"] if precision == "f32" else g["whole_pts"]
    np.testing.assert_array_equal(cloud.reshape(-1, 3).cpu().numpy(), want)
    crop = pkg.data_utils.get_point_cloud_from_depth(raw, K, bbox=[10, 50, 20, 90], precision=precision)
    np.testing.assert_array_equal(crop.cpu().numpy(), want.reshape(240, 320, 3)[10:50, 20:90])


def _scene(seed, H, W, n_det):
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    depth = 900.0 + 0.3 * xx + 0.1 * yy + rs.normal(0, 2.0, (H, W))
    masks = np.zeros((n_det, H, W), np.uint8)
    for i in range(n_det):
        cy, cx = rs.randint(0, H), rs.randint(0, W)
        ry, rx = rs.randint(4, H // 3), rs.randint(4, W // 3)
        e = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2
        depth[e < 0.7] = 500.0 + 10 * i + 60.0 * e[e < 0.7]
        masks[i] = e < 1.0
    masks[n_det - 1] = 0
    masks[n_det - 1, 5:9, 7:14] = 1                                      # 28 pixels: dropped
    depth[rs.rand(H, W) < 0.05] = 0.0
    return (np.round(depth).astype(np.float32) * 0.5 / 1000.0).astype(np.float32), masks, rs.rand(n_det)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["f32", "f64"])
@pytest.mark.parametrize("shape", [(480, 640, 7), (97, 131, 5)])
def test_gpu_matches_oracle(pkg, oracle, dev, precision, shape):
    H, W, D = shape
    depth, masks, scores = _scene(H + D, H, W, D)
    K = np.array([[572.4114, 0, 325.2611], [0, 573.57043, 242.04899], [0, 0, 1]])
    radius = 0.21
    pts, rc, sc, kept, _ = oracle.prepare_instances(depth, K, masks, scores, radius, NS, IMG,
                                                    np.random.RandomState(5), precision)
    out = _run(pkg, dev, depth, K, masks, scores, radius, precision, 5)
    assert out["kept"] == kept and (D - 1) not in kept
    np.testing.assert_array_equal(out["pts"].cpu().numpy(), pts)
    np.testing.assert_array_equal(out["rgb_choose"].cpu().numpy(), rc)
    np.testing.assert_array_equal(out["score"].cpu().numpy(), sc)
    # the draw covers both branches of the reference (with / without replacement)
    assert (out["n_candidates"] > NS).any() or H < 200


@pytest.mark.gpu
def test_gpu_rejects_bad_arguments(pkg, dev):
    depth = torch.ones((8, 8), device=dev)
    with pytest.raises(TypeError):
        pkg.data_utils.prepare_instances(depth, np.eye(3), torch.ones((1, 8, 8), device=dev), [1.0], 0.1)
    with pytest.raises(ValueError):
        pkg.data_utils.prepare_instances(depth, np.eye(3), torch.ones((1, 9, 8), dtype=torch.uint8, device=dev),
                                         [1.0], 0.1)
    with pytest.raises(RuntimeError):  # nothing survives: 8x8 ones has a mask but a tiny radius drops every point
        pkg.data_utils.prepare_instances(depth * 0, np.eye(3), torch.ones((1, 8, 8), dtype=torch.uint8, device=dev),
                                         [1.0], 0.1)
