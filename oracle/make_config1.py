"""oracle/make_config1.py -- BASELINE.json configs[0]: the PEM geometric stage on the bundled
Data/Example (obj_000005.ply + depth.png), 1 instance, random-init weights, run through the
REFERENCE's own modules on CPU (this container only).  TEST INFRASTRUCTURE.

Writes tests/golden/config1_example.npz: every tensor the reference hands to / gets back from
its geometric functions during steps 2-4 of Net.forward (model/pose_estimation_model.py:29-55),
recorded by wrapping those functions (nothing in the reference is modified).  The ViT
(step 1, needs `timm`) is replaced by seeded N(0,1) dense features, as SURVEY.md 8d prescribes.
"""
import os
import sys
import types

import cv2
import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_live  # noqa: E402

EX = "/root/reference/SAM-6D/Data/Example"
PEM = ref_live.REF_PEM


def load_ply(path):
    with open(path) as f:
        lines = f.read().split("\n")
    nv = int([l for l in lines if l.startswith("element vertex")][0].split()[-1])
    nf = int([l for l in lines if l.startswith("element face")][0].split()[-1])
    h = lines.index("end_header") + 1
    v = np.array([l.split()[:3] for l in lines[h:h + nv]], dtype=np.float64)
    f = np.array([l.split()[1:4] for l in lines[h + nv:h + nv + nf]], dtype=np.int64)
    return v, f


def sample_surface(v, f, n, rs):
    a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    area = 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1)
    tri = rs.choice(len(f), size=n, p=area / area.sum())
    u, w = rs.rand(n, 1), rs.rand(n, 1)
    flip = (u + w) > 1
    u, w = np.where(flip, 1 - u, u), np.where(flip, 1 - w, w)
    return a[tri] + u * (b[tri] - a[tri]) + w * (c[tri] - a[tri])


def cfg_obj(d):
    o = types.SimpleNamespace()
    for k, v in d.items():
        setattr(o, k, cfg_obj(v) if isinstance(v, dict) else v)
    return o


def main():
    mu, pu, ext = ref_live.load()
    import coarse_point_matching as cpm
    import fine_point_matching as fpm
    import transformer as tfm
    rs = np.random.RandomState(1)
    torch.manual_seed(1)

    # CAD model (mm -> m), template cloud = FPS of a dense surface sample (feature_extraction.py:152-158)
    v, f = load_ply(os.path.join(EX, "obj_000005.ply"))
    model = sample_surface(v, f, 1024, rs).astype(np.float32) / 1000.0
    cand = torch.from_numpy(sample_surface(v, f, 8192, rs).astype(np.float32) / 1000.0)[None]
    idx = ext.furthest_point_sampling(cand, 2048).long()
    dense_po = cand[0, idx[0]][None].contiguous()
    radius = dense_po.norm(dim=2).max(dim=1)[0]                       # feature_extraction.py:132-133
    # observed cloud: depth back-projection (data_utils.py:92-110), points near an object-sized region
    depth = cv2.imread(os.path.join(EX, "depth.png"), cv2.IMREAD_UNCHANGED).astype(np.float32) / 1000.0
    K = np.array(yaml.safe_load(open(os.path.join(EX, "camera.json")))["cam_K"], dtype=np.float32).reshape(3, 3)
    H, W = depth.shape
    xm, ym = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    cloud = np.stack([(xm - K[0, 2]) * depth / K[0, 0], (ym - K[1, 2]) * depth / K[1, 1], depth], -1).reshape(-1, 3)
    cloud = cloud[cloud[:, 2] > 0]
    centre = cloud[np.argmin(np.abs(cloud[:, 0]) + np.abs(cloud[:, 1]))]   # valid pixel nearest the optical axis
    d = np.linalg.norm(cloud - centre, axis=1)
    near = np.nonzero(d < 1.2 * radius.item())[0]                          # run_inference_custom_pytorch.py:329-342
    choose = rs.choice(near, 2048, replace=len(near) <= 2048)
    pts = torch.from_numpy(cloud[choose])[None]
    dense_pm = (pts / radius.reshape(-1, 1, 1)).contiguous()
    dense_po_n = (dense_po / radius.reshape(-1, 1, 1)).contiguous()
    dense_fm, dense_fo = torch.randn(1, 2048, 256), torch.randn(1, 2048, 256)   # stand-in for the ViT
    model_t = torch.from_numpy(model)[None]

    cfg = cfg_obj(yaml.safe_load(open(os.path.join(PEM, "config", "base.yaml")))["model"])
    geo = tfm.GeometricStructureEmbedding(cfg.geo_embedding).eval()
    coarse = cpm.CoarsePointMatching(cfg.coarse_point_matching).eval()
    fine = fpm.FinePointMatching(cfg.fine_point_matching).eval()

    rec = {}

    def wrap(mod, name, tag, keep_in, keep_out):
        orig = getattr(mod, name)

        def w(*a, **k):
            out = orig(*a, **k)
            outs = out if isinstance(out, tuple) else (out,)
            n = sum(1 for kk in rec if kk.startswith(tag + "_call"))
            for i in keep_in:
                rec["%s_call%d_in%d" % (tag, n, i)] = a[i].detach().clone()
            for i in keep_out:
                rec["%s_call%d_out%d" % (tag, n, i)] = outs[i].detach().clone()
            rec["%s_call%d" % (tag, n)] = torch.zeros(1)
            return out
        setattr(mod, name, w)

    wrap(cpm, "compute_feature_similarity", "csim", (0, 1), (0,))
    wrap(cpm, "compute_coarse_Rt", "coarse", (0, 1, 2, 3), (0, 1))
    wrap(fpm, "compute_feature_similarity", "fsim", (0, 1), ())
    wrap(fpm, "compute_fine_Rt", "fine", (1, 2, 3), (0, 1, 2))

    with torch.no_grad():
        bg_point = torch.ones(1, 1, 3) * 100
        spm, sfm, idx_m = mu.sample_pts_feats(dense_pm, dense_fm, 196, return_index=True)
        geo_m = geo(torch.cat([bg_point, spm], dim=1))
        spo, sfo, idx_o = mu.sample_pts_feats(dense_po_n, dense_fo, 196, return_index=True)
        geo_o = geo(torch.cat([bg_point, spo], dim=1))
        torch.manual_seed(5)
        state = torch.get_rng_state()
        init_R, init_t = coarse(spm, sfm, geo_m, spo, sfo, geo_o, radius, model_t)
        torch.set_rng_state(state)
        rand = torch.rand(1, 18000)
        pred_R, pred_t, score = fine(dense_pm, dense_fm, geo_m, idx_m, dense_po_n, dense_fo, geo_o, idx_o,
                                     radius, model_t, init_R, init_t)
    out = {
        "dense_pm": dense_pm, "dense_po": dense_po_n,
        # dense_fm / dense_fo are torch.manual_seed(1); randn(1,2048,256) x2 -- regenerated by the
        # tests and verified against these checksums instead of being stored (4 MB)
        "dense_f_checksum": torch.stack([dense_fm.double().sum(), dense_fo.double().sum(),
                                         dense_fm[0, 7, 3].double(), dense_fo[0, 2047, 255].double()]),
        "sparse_fm": sfm.contiguous(), "sparse_fo": sfo.contiguous(),
        "radius": radius, "model_pts": rec["coarse_call0_in3"], "rand": rand,
        "sparse_pm": spm.contiguous(), "sparse_po": spo.contiguous(), "fps_m": idx_m, "fps_o": idx_o,
        "coarse_f1": rec["csim_call0_in0"], "coarse_f2": rec["csim_call0_in1"], "coarse_atten": rec["csim_call0_out0"],
        "init_R": rec["coarse_call0_out0"], "init_t": rec["coarse_call0_out1"],
        "fine_f1": rec["fsim_call0_in0"], "fine_f2": rec["fsim_call0_in1"],
        "fine_R": rec["fine_call0_out0"], "fine_t": rec["fine_call0_out1"], "fine_score": rec["fine_call0_out2"],
        "pred_t_scaled": pred_t,
    }
    assert torch.equal(rec["coarse_call0_in0"], rec["csim_call0_out0"])
    path = os.path.join(ROOT, "tests", "golden", "config1_example.npz")
    np.savez_compressed(path, **{k: v.numpy() for k, v in out.items()})
    print("wrote", path, os.path.getsize(path), "bytes; near pts", len(near), "radius", radius.item(),
          "init_R det", torch.det(init_R).item(), "score", score.item())


if __name__ == "__main__":
    main()
