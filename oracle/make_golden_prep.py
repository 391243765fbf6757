"""oracle/make_golden_prep.py -- golden vectors for the host-side instance preparation (SURVEY.md
section 8f, rank 4) from the LIVE reference in this container.

TEST INFRASTRUCTURE ONLY.  Imports the unmodified run_inference_custom_pytorch.py and calls its
get_test_data on a synthetic RGB-D scene written to a temporary directory.  Modules this image lacks
are stubbed at the file-format edge only -- imageio.imread (reads .npy), pycocotools.mask (the masks
travel as arrays instead of RLE strings) and trimesh.load_mesh (seeded model points) -- so every line
of the reference between "decoded mask" and "input_data" runs as written.  numpy here is 2.3, so the
recorded outputs carry the float64 promotion of NEP 50 (the oracle's "f64" mode); the reference's
pinned numpy 1.26 ("f32" mode) is pinned on get_point_cloud_from_depth with float32 intrinsics.
Writes tests/golden/instance_prep.npz.
"""
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_live  # noqa: E402

H, W = 240, 320
N_SAMPLE, N_MODEL, IMG_SIZE = 2048, 1024, 224
K = [300.0, 0.0, 159.5, 0.0, 301.5, 119.5, 0.0, 0.0, 1.0]


def make_scene(seed=7):
    """Tilted background plane ~1.1 m with holes, six blobs in front of it; masks are ellipses a bit
    larger than the blobs (so the radius filter has background points to drop), one touches the image
    border, one is tiny (<= 32 px: dropped), one is large (> N_SAMPLE points: drawn without replacement)."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    depth = 1100.0 + 0.35 * xx - 0.2 * yy + rng.normal(0, 1.5, (H, W))
    blobs = [(60, 70, 38, 30), (150, 230, 46, 52), (20, 300, 30, 30), (200, 40, 9, 9), (120, 150, 3, 3), (215, 300, 36, 30)]
    masks = []
    for cy, cx, ry, rx in blobs:
        e = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2
        inner = e < 0.8
        depth[inner] = 640.0 + 55.0 * e[inner] + rng.normal(0, 0.8, inner.sum())
        masks.append((e < 1.0).astype(np.uint8))
    depth[rng.rand(H, W) < 0.03] = 0.0  # sensor holes
    depth = np.clip(np.round(depth), 0, 65535).astype(np.uint16)
    rgb = rng.randint(0, 256, (H, W, 3)).astype(np.uint8)
    scores = np.asarray([0.9, 0.8, 0.7, 0.6, 0.5, 0.1 + 0.3], np.float64)
    return depth, rgb, np.stack(masks), scores


def install_stubs(mask_table, model_mm):
    imageio = types.ModuleType("imageio")
    imageio.imread = lambda path: np.load(path)
    sys.modules["imageio"] = imageio
    coco = types.ModuleType("pycocotools")
    cm = types.ModuleType("pycocotools.mask")

    def fr_py_objects(seg, h, w):
        raise ValueError("already an RLE")  # the reference falls back to rle = seg

    cm.frPyObjects = fr_py_objects
    cm.decode = lambda rle: np.asfortranarray(mask_table[rle["counts"]])  # pycocotools returns F order
    coco.mask = cm
    sys.modules["pycocotools"], sys.modules["pycocotools.mask"] = coco, cm
    tm = types.ModuleType("trimesh")

    class _Mesh:
        def sample(self, n):
            return model_mm[:n].astype(np.float64)

    tm.load_mesh = lambda path: _Mesh()
    sys.modules["trimesh"] = tm


def main():
    depth, rgb, masks, scores = make_scene()
    model_mm = np.random.RandomState(3).normal(0, 45.0, (N_MODEL, 3))
    install_stubs({i: masks[i] for i in range(len(masks))}, model_mm)
    ref_live.load()
    sys.path.insert(0, ref_live.REF_PEM)
    argv, sys.argv = sys.argv, sys.argv[:1]
    import run_inference_custom_pytorch as ref  # the reference script, unmodified (main is guarded)
    sys.argv = argv
    from utils import data_utils

    class Cfg:
        n_sample_model_point, n_sample_observed_point, img_size, rgb_mask_flag = N_MODEL, N_SAMPLE, IMG_SIZE, True

    with tempfile.TemporaryDirectory() as d:
        np.save(os.path.join(d, "rgb.npy"), rgb)
        np.save(os.path.join(d, "depth.npy"), depth)
        json.dump({"cam_K": K, "depth_scale": 1.0}, open(os.path.join(d, "cam.json"), "w"))
        dets = [{"scene_id": 0, "image_id": 0, "category_id": 1, "bbox": [0, 0, 1, 1], "score": float(scores[i]),
                 "segmentation": {"size": [H, W], "counts": i}} for i in range(len(masks))]
        json.dump(dets, open(os.path.join(d, "seg.json"), "w"))
        np.random.seed(11)
        data, img, whole_pts, model_points, kept = ref.get_test_data(
            os.path.join(d, "rgb.npy"), os.path.join(d, "depth.npy"), os.path.join(d, "cam.json"), "unused.ply",
            os.path.join(d, "seg.json"), 0.2, Cfg(), "cpu")
    out = {
        "depth_u16": depth, "masks": masks, "scores": scores, "K": np.asarray(K), "model_mm": model_mm,
        "pts": data["pts"].numpy(), "rgb_choose": data["rgb_choose"].numpy(), "score": data["score"].numpy(),
        "model": data["model"][0].numpy(), "kept": np.asarray([d_["segmentation"]["counts"] for d_ in kept]),
        "whole_pts": whole_pts.astype(np.float32),
    }
    # numpy-1.26 arithmetic of the back-projection: the reference function on float32 intrinsics
    depth_m = depth.astype(np.float32) * 1.0 / 1000.0
    out["whole_pts_f32"] = data_utils.get_point_cloud_from_depth(depth_m, np.asarray(K, np.float32).reshape(3, 3)).reshape(-1, 3)
    assert out["whole_pts_f32"].dtype == np.float32
    # get_bbox / get_resize_rgb_choose on their own (edge boxes)
    m = np.logical_and(masks > 0, depth_m[None] > 0)
    out["bboxes"] = np.asarray([data_utils.get_bbox(m[i]) for i in range(len(masks))])
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "instance_prep.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), {k: v.shape for k, v in out.items()}, "kept", out["kept"])


if __name__ == "__main__":
    main()
