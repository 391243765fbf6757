"""The drop-in boundary exercised ON THE GPU through the reference's own, unmodified call sites
(model/pointnet2/pointnet2_utils.py:25-33,78,118,255,290; utils/model_utils.py:70-84;
model/coarse_point_matching.py:45-62; model/fine_point_matching.py:62-77).

The reference's Python files are not part of this repository: oracle/build_ref.snapshot_reference_py() copies the
handful of call-site files into the git-ignored oracle/_ref/pem_py at build time (test infrastructure; it travels
to the GPU box like oracle/_ref's compiled extension).  Here they are imported with this package's module
registered as `pointnet2._ext` (dropin.install_pointnet2_ext) and the geometric functions swapped in
(dropin.patch_model_utils / patch_call_sites); everything else -- QueryAndGroup, sample_pts_feats,
GeometricStructureEmbedding, the transformers, PositionalEncoding, CoarsePointMatching.forward,
FinePointMatching.forward -- is the reference's code running on CUDA tensors.

Network layers on the GPU differ from the CPU run that wrote the golden file by float noise, so results are
compared where the drop-in sits: every call the reference makes into a replaced function is recorded and
replayed through the oracle on the SAME inputs."""
import os
import sys
import types

import numpy as np
import pytest
import torch
import yaml

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu


def _cfg(d):
    o = types.SimpleNamespace()
    for k, v in d.items():
        setattr(o, k, _cfg(v) if isinstance(v, dict) else v)
    return o


@pytest.fixture(scope="module")
def ref(pkg):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_live
    if ref_live.pem_root() is None:
        pytest.skip("no reference call-site snapshot under oracle/_ref/pem_py (run __graft_entry__.build() where "
                    "/root/reference exists)")
    ext = pkg.install_pointnet2_ext()
    mods = ref_live.load_call_sites(ext)
    yield mods
    for name in ("pointnet2_utils", "model_utils", "transformer", "coarse_point_matching", "fine_point_matching",
                 "pytorch_utils", "loss_utils"):
        sys.modules.pop(name, None)


def test_reference_pointnet2_utils_runs_on_the_dropin(pkg, oracle, dev, ref):
    pu = ref["pointnet2_utils"]
    assert pu._ext is pkg.pointnet2_ext                      # the reference imported OUR module as pointnet2._ext
    b = pkg.synth.make_batch([50, 51], n_dense=1024, n_model=64, feat_dim=32, n_proposal1=8)
    xyz = b["dense_po"].to(dev)
    feat = b["dense_fo"].transpose(1, 2).contiguous().to(dev)
    # FurthestPointSampling / GatherOperation through the reference's autograd Functions
    # (the reference passes the count as the LENGTH of a tensor: utils/model_utils.py:75, pointnet2_utils.py:77)
    idx = pu.furthest_point_sample(xyz, torch.ones(196, dtype=torch.int32))
    want = oracle.fps(b["dense_po"].numpy(), 196)
    assert np.array_equal(idx.cpu().numpy(), want)
    g = pu.gather_operation(feat, idx)
    assert np.array_equal(g.cpu().numpy(), oracle.gather(feat.cpu().numpy(), want))
    # QueryAndGroup (the reference's nn.Module: ball_query + 2 x grouping_operation + subtract + cat)
    for r, ns in ((0.1, 32), (0.2, 64)):
        qg = pu.QueryAndGroup(r, ns, use_xyz=True)
        out = qg(xyz, xyz + 1e-8, xyz.transpose(1, 2).contiguous())
        wq, _ = oracle.query_and_group(b["dense_po"].numpy(), (b["dense_po"] + 1e-8).numpy(),
                                       b["dense_po"].transpose(1, 2).contiguous().numpy(), r, ns)
        assert np.array_equal(out.cpu().numpy(), wq)
    # the reference's errors for bad tensors still surface as RuntimeError (utils.h:20-35)
    with pytest.raises(RuntimeError):
        pu.furthest_point_sample(xyz.double(), torch.ones(8, dtype=torch.int32))


def test_reference_sample_pts_feats_unpatched_on_config1(pkg, dev, ref):
    """utils/model_utils.sample_pts_feats as written (transposes + gather_operation), on our `_ext`."""
    mu = ref["model_utils"]
    g = np.load(os.path.join(GOLDEN, "config1_example.npz"))
    torch.manual_seed(1)
    fm, fo = torch.randn(1, 2048, 256), torch.randn(1, 2048, 256)
    for pk, f, ik, sk, sfk in (("dense_pm", fm, "fps_m", "sparse_pm", "sparse_fm"),
                               ("dense_po", fo, "fps_o", "sparse_po", "sparse_fo")):
        p, sf, idx = mu.sample_pts_feats(torch.from_numpy(g[pk]).to(dev), f.to(dev), 196, return_index=True)
        assert np.array_equal(idx.cpu().numpy(), g[ik])
        assert np.array_equal(p.cpu().numpy(), g[sk]) and np.array_equal(sf.cpu().numpy(), g[sfk])


def test_reference_matching_modules_on_the_dropin(pkg, oracle, dev, ref):
    """CoarsePointMatching.forward + FinePointMatching.forward of the reference (random-init, config-1 inputs) on
    CUDA with the geometric functions swapped in; every call into a swapped function is replayed by the oracle."""
    mu, cpm, fpm, tfm = ref["model_utils"], ref["coarse_point_matching"], ref["fine_point_matching"], ref["transformer"]
    fast = pkg.model_utils
    g = np.load(os.path.join(GOLDEN, "config1_example.npz"))
    c1 = {k: torch.from_numpy(np.ascontiguousarray(g[k])) for k in g.files}
    torch.manual_seed(1)
    dense_fm, dense_fo = torch.randn(1, 2048, 256), torch.randn(1, 2048, 256)
    cfg = _cfg(yaml.safe_load(open(os.path.join(ref["root"], "config", "base.yaml")))["model"])
    geo = tfm.GeometricStructureEmbedding(cfg.geo_embedding).eval().to(dev)
    coarse = cpm.CoarsePointMatching(cfg.coarse_point_matching).eval().to(dev)
    fine = fpm.FinePointMatching(cfg.fine_point_matching).eval().to(dev)

    calls = []
    rand = c1["rand"].to(dev)

    def rec(name, fn, **fixed):
        def w(*a, **k):
            out = fn(*a, **k, **fixed)
            calls.append((name, [x.detach().cpu() if isinstance(x, torch.Tensor) else x for x in a],
                          [x.detach().cpu() for x in (out if isinstance(out, tuple) else (out,))]))
            return out
        return w

    pkg.dropin.patch_model_utils(mu)
    pkg.dropin.patch_call_sites(cpm, fpm)
    assert cpm.compute_coarse_Rt is fast.compute_coarse_Rt and fpm.compute_fine_Rt is fast.compute_fine_Rt
    cpm.compute_feature_similarity = rec("coarse_sim", fast.compute_feature_similarity)
    cpm.compute_coarse_Rt = rec("coarse_Rt", fast.compute_coarse_Rt, rand=rand)
    fpm.compute_feature_similarity = rec("fine_sim", fast.compute_feature_similarity)
    fpm.compute_fine_Rt = rec("fine_Rt", fast.compute_fine_Rt)

    d = lambda k: c1[k].to(dev)
    with torch.no_grad():
        bg_point = torch.ones(1, 1, 3, device=dev) * 100
        spm, sfm, idx_m = mu.sample_pts_feats(d("dense_pm"), dense_fm.to(dev), 196, return_index=True)
        spo, sfo, idx_o = mu.sample_pts_feats(d("dense_po"), dense_fo.to(dev), 196, return_index=True)
        assert torch.equal(idx_m.cpu(), c1["fps_m"]) and torch.equal(idx_o.cpu(), c1["fps_o"])
        geo_m = geo(torch.cat([bg_point, spm], dim=1))
        geo_o = geo(torch.cat([bg_point, spo], dim=1))
        radius, model_t = d("radius"), (c1["model_pts"] * (c1["radius"].reshape(-1, 1, 1) + 1e-6)).to(dev)
        init_R, init_t = coarse(spm, sfm, geo_m, spo, sfo, geo_o, radius, model_t)
        pred_R, pred_t, score = fine(d("dense_pm"), dense_fm.to(dev), geo_m, idx_m, d("dense_po"), dense_fo.to(dev),
                                     geo_o, idx_o, radius, model_t, init_R, init_t)
    torch.cuda.synchronize()
    assert pred_R.shape == (1, 3, 3) and pred_t.shape == (1, 3) and torch.isfinite(pred_R).all()
    assert (torch.det(pred_R.cpu()) - 1).abs().max().item() < 1e-4
    names = [c[0] for c in calls]
    assert names.count("coarse_sim") >= 1 and names.count("coarse_Rt") == 1 and names.count("fine_Rt") == 1
    for name, a, out in calls:
        if name in ("coarse_sim", "fine_sim"):
            want = oracle.feature_similarity(a[0], a[1], a[3])
            assert (out[0] - want).abs().max().item() < 3e-5, name
        elif name == "coarse_Rt":
            atten, p1, p2, model_n = a[0], a[1], a[2], a[3]
            oR, ot, ost = oracle.coarse_Rt(atten, p1, p2, model_n, a[4], a[5], c1["rand"])
            # our winner, under the oracle's scorer, is at least as good as the best reproducible (well-posed)
            # hypothesis of the oracle's list; identical when the oracle's winner is well posed
            Htop = torch.gather(ost["H"], 1, ost["top"].reshape(1, -1, 1, 1).expand(-1, -1, 3, 3))
            sv = torch.linalg.svdvals(Htop.double())
            posed = sv[..., 1] / sv[..., 0].clamp_min(1e-30) > 1e-2
            q = oracle.score_hypotheses(out[0][:, None], out[1][:, None], p1, ost["w1"], model_n)[:, 0]
            best_posed = torch.where(posed, ost["scores"], torch.zeros(())).max(1)[0]
            assert (q >= best_posed * (1 - 1e-3)).all(), (q, best_posed)
            if posed[0, ost["scores"][0].argmax()]:
                assert oracle.rotation_angle_deg(out[0], oR).max().item() < 0.1
        elif name == "fine_Rt":
            oR, ot, os_, _ = oracle.fine_Rt(a[0], a[1], a[2], a[3])
            assert oracle.rotation_angle_deg(out[0], oR).max().item() < 0.1
            assert (out[1] - ot).abs().max().item() < 1e-4
            assert (out[2] - os_).abs().max().item() < 5e-3
    # the caller's own post-processing ran on our outputs (fine_point_matching.py:78: t *= radius + 1e-6)
    fr = [c for c in calls if c[0] == "fine_Rt"][0]
    assert torch.allclose(pred_t.cpu(), fr[2][1] * (c1["radius"].reshape(-1, 1) + 1e-6), atol=1e-7)
