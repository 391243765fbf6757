"""oracle/make_golden.py -- write the committed fixtures under tests/golden/ from the LIVE,
UNMODIFIED reference (/root/reference, this container only).  TEST INFRASTRUCTURE.

Every array in the fixtures is an output of the reference's own code
(pointnet2 CPU extension built by oracle/build_ref.py; utils/model_utils.py imported through
oracle/ref_live.py) on seeded inputs that are stored alongside.  The only deviation from a
plain call is that the uniform draw of compute_coarse_Rt (torch.rand inside MU:293) is
reproduced by seeding torch's CPU generator immediately before the call and stored as `rand`.

Run:  python oracle/make_golden.py        (needs /root/reference)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_live  # noqa: E402
import __graft_entry__ as entry  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    mu, pu, ext = ref_live.load()
    pkg = entry.load_package()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(32)
    np.random.seed(324)  # the seeds of the reference's own op scripts (ov_test_ball_query.py:167-168)

    # ---- pointnet2 ops, small shapes with the edge cases the CPU loops special-case ----------
    g = torch.Generator().manual_seed(1)
    xyz = torch.randn(3, 600, 3, generator=g) * 0.35
    xyz[:, 11] = 0.01          # near origin: skipped by the CPU FPS (sampling.cpp:102-103)
    xyz[:, 50:60] = xyz[:, 5:15]  # exact duplicates: distance ties
    fps_idx = ext.furthest_point_sampling(xyz, 96)
    feat = torch.randn(3, 8, 600, generator=g)
    gidx = fps_idx.clone()
    gidx[0, 0] = -1            # invalid -> 0.0 (sampling.cpp:34-40)
    gathered = ext.gather_points(feat, gidx)
    centres = (xyz[:, :128] + 1e-8).contiguous()
    centres[:, -1] = 40.0      # no neighbour: all-zero row (ball_query.cpp:75-77)
    bq = {}
    for r, ns in ((0.1, 32), (0.2, 64)):
        bq[(r, ns)] = ext.ball_query(centres, xyz, r, ns)
    grouped = ext.group_points(feat, bq[(0.2, 64)])
    qg = pu.QueryAndGroup(0.1, 32)(xyz, (xyz + 1e-8).contiguous(), xyz.transpose(1, 2).contiguous())
    np.savez_compressed(
        os.path.join(OUT, "pointnet2_small.npz"),
        xyz=xyz.numpy(), fps_idx=fps_idx.numpy(), feat=feat.numpy(), gather_idx=gidx.numpy(),
        gathered=gathered.numpy(), centres=centres.numpy(), bq_r01_ns32=bq[(0.1, 32)].numpy().astype(np.int16),
        bq_r02_ns64=bq[(0.2, 64)].numpy().astype(np.int16), grouped=grouped.numpy(), qg_r01_ns32=qg.numpy())

    # ---- default shapes: FPS 2048 -> 196 and PE ball queries on a synthetic instance ----------
    b = pkg.synth.make_batch([1])
    idx_m = ext.furthest_point_sampling(b["dense_pm"], 196)
    idx_o = ext.furthest_point_sampling(b["dense_po"], 196)
    po = b["dense_po"]
    bq1 = ext.ball_query((po + 1e-8).contiguous(), po, 0.1, 32)
    bq2 = ext.ball_query((po + 1e-8).contiguous(), po, 0.2, 64)
    np.savez_compressed(
        os.path.join(OUT, "pointnet2_default.npz"),
        dense_pm=b["dense_pm"].numpy(), dense_po=po.numpy(), fps_m=idx_m.numpy().astype(np.int16),
        fps_o=idx_o.numpy().astype(np.int16), bq_r01_ns32=bq1.numpy().astype(np.int16),
        bq_r02_ns64=bq2.numpy().astype(np.int16))

    # ---- coarse stage, reduced sizes (2 instances) ------------------------------------------
    n1 = 48
    f1 = torch.randn(2, n1 + 1, 32, generator=g)
    f2 = f1[:, torch.randperm(n1 + 1, generator=g)] + 0.3 * torch.randn(2, n1 + 1, 32, generator=g)
    atten = mu.compute_feature_similarity(f1, f2, "cosine", 0.1, True)
    p1 = torch.randn(2, n1, 3, generator=g) * 0.3
    p2 = torch.randn(2, n1, 3, generator=g) * 0.3
    model = torch.randn(2, 200, 3, generator=g) * 0.3
    torch.manual_seed(7)
    state = torch.get_rng_state()
    R, t = mu.compute_coarse_Rt(atten, p1, p2, model, 500, 40)
    torch.set_rng_state(state)
    rand = torch.rand(2, 1500)
    np.savez_compressed(os.path.join(OUT, "coarse_small.npz"), f1=f1.numpy(), f2=f2.numpy(), atten=atten.numpy(),
                        pts1=p1.numpy(), pts2=p2.numpy(), model=model.numpy(), rand=rand.numpy(), R=R.numpy(),
                        t=t.numpy(), n_proposal1=500, n_proposal2=40)

    # ---- coarse stage, default sizes (1 instance) -------------------------------------------
    import pem_oracle as O
    sp, sf, _ = O.sample_pts_feats(b["dense_pm"].numpy(), b["dense_fm"].numpy(), 196)
    tp, tf, _ = O.sample_pts_feats(b["dense_po"].numpy(), b["dense_fo"].numpy(), 196)
    sp, sf, tp, tf = (torch.from_numpy(np.ascontiguousarray(x)) for x in (sp, sf, tp, tf))
    bg = b["bg"][:, None, :]
    attc = mu.compute_feature_similarity(torch.cat([bg, sf], 1), torch.cat([bg, tf], 1), "cosine", 0.1, True)
    torch.manual_seed(11)
    state = torch.get_rng_state()
    Rc, tc = mu.compute_coarse_Rt(attc, sp, tp, b["model"], 6000, 300)
    torch.set_rng_state(state)
    randc = torch.rand(1, 18000)
    np.savez_compressed(os.path.join(OUT, "coarse_default.npz"), atten=attc.numpy(), pts1=sp.numpy(), pts2=tp.numpy(),
                        model=b["model"].numpy(), rand=randc.numpy(), R=Rc.numpy(), t=tc.numpy())

    # ---- fine stage, reduced size -----------------------------------------------------------
    bs = pkg.synth.make_batch([2, 3], n_dense=384, n_model=256, feat_dim=64, n_proposal1=16)
    bgs = bs["bg"][:, None, :]
    ff1, ff2 = torch.cat([bgs, bs["dense_fm"]], 1), torch.cat([bgs, bs["dense_fo"]], 1)
    attf = mu.compute_feature_similarity(ff1, ff2, "cosine", 0.1, True)
    Rf, tf_, sf_ = mu.compute_fine_Rt(attf, bs["dense_pm"], bs["dense_po"], bs["model"])
    np.savez_compressed(os.path.join(OUT, "fine_small.npz"), f1=ff1.numpy(), f2=ff2.numpy(),
                        pts1=bs["dense_pm"].numpy(), pts2=bs["dense_po"].numpy(), model=bs["model"].numpy(),
                        R=Rf.numpy(), t=tf_.numpy(), score=sf_.numpy())

    # ---- weighted Procrustes ----------------------------------------------------------------
    src = torch.randn(4, 50, 3, generator=g)
    ref = torch.randn(4, 50, 3, generator=g)
    w = torch.rand(4, 50, generator=g)
    Rw, tw = mu.weighted_procrustes(src, ref, w, weight_thresh=0.3)
    np.savez_compressed(os.path.join(OUT, "procrustes.npz"), src=src.numpy(), ref=ref.numpy(), w=w.numpy(),
                        R=Rw.numpy(), t=tw.numpy(), weight_thresh=0.3)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
