"""GPU parity against the committed golden vectors (tests/golden/*.npz), which are outputs of
the live, unmodified reference written by oracle/make_golden.py."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from test_dense_gpu import check_coarse_against_oracle

pytestmark = pytest.mark.gpu


def _g(name):
    return np.load(os.path.join(GOLDEN, name))


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_pointnet2_small_golden(pkg, dev):
    g, ext = _g("pointnet2_small.npz"), pkg.pointnet2_ext
    xyz, feat = _t(g["xyz"], dev), _t(g["feat"], dev)
    np.testing.assert_array_equal(ext.furthest_point_sampling(xyz, 96).cpu().numpy(), g["fps_idx"])
    np.testing.assert_array_equal(ext.gather_points(feat, _t(g["gather_idx"], dev)).cpu().numpy(), g["gathered"])
    cen = _t(g["centres"], dev)
    for key, r, ns in (("bq_r01_ns32", 0.1, 32), ("bq_r02_ns64", 0.2, 64)):
        np.testing.assert_array_equal(ext.ball_query(cen, xyz, r, ns).cpu().numpy(), g[key].astype(np.int32))
    idx = _t(g["bq_r02_ns64"].astype(np.int32), dev)
    np.testing.assert_array_equal(ext.group_points(feat, idx).cpu().numpy(), g["grouped"])
    qg = pkg.pointnet2_utils.QueryAndGroup(0.1, 32)(xyz, (xyz + 1e-8).contiguous(), xyz.transpose(1, 2).contiguous())
    np.testing.assert_array_equal(qg.cpu().numpy(), g["qg_r01_ns32"])


def test_pointnet2_default_golden(pkg, dev):
    g, ext = _g("pointnet2_default.npz"), pkg.pointnet2_ext
    pm, po = _t(g["dense_pm"], dev), _t(g["dense_po"], dev)
    np.testing.assert_array_equal(ext.furthest_point_sampling(pm, 196).cpu().numpy(), g["fps_m"].astype(np.int32))
    np.testing.assert_array_equal(ext.furthest_point_sampling(po, 196).cpu().numpy(), g["fps_o"].astype(np.int32))
    cen = (po + 1e-8).contiguous()
    np.testing.assert_array_equal(ext.ball_query(cen, po, 0.1, 32).cpu().numpy(), g["bq_r01_ns32"].astype(np.int32))
    np.testing.assert_array_equal(ext.ball_query(cen, po, 0.2, 64).cpu().numpy(), g["bq_r02_ns64"].astype(np.int32))


def test_coarse_default_golden(pkg, oracle, dev):
    g = _g("coarse_default.npz")
    oR, ot = check_coarse_against_oracle(pkg, oracle, dev, g["atten"], g["pts1"], g["pts2"], g["model"], 6000, 300,
                                         g["rand"], min_common=5)
    # the oracle reproduces the stored reference output bit for bit (also checked on CPU)
    assert torch.equal(oR, torch.from_numpy(g["R"])) and torch.equal(ot, torch.from_numpy(g["t"]))


def test_coarse_small_golden(pkg, oracle, dev):
    g = _g("coarse_small.npz")
    att = pkg.model_utils.compute_feature_similarity(_t(g["f1"], dev), _t(g["f2"], dev), "cosine", 0.1, True)
    assert (att.cpu() - torch.from_numpy(g["atten"])).abs().max().item() < 3e-5
    check_coarse_against_oracle(pkg, oracle, dev, g["atten"], g["pts1"], g["pts2"], g["model"],
                                int(g["n_proposal1"]), int(g["n_proposal2"]), g["rand"], min_common=5)


@pytest.mark.parametrize("path", ["materialised", "fused"])
def test_fine_small_golden(pkg, oracle, dev, path):
    g, mu = _g("fine_small.npz"), pkg.model_utils
    f1, f2 = _t(g["f1"], dev), _t(g["f2"], dev)
    p1, p2, model = _t(g["pts1"], dev), _t(g["pts2"], dev), _t(g["model"], dev)
    if path == "fused":
        R, t, s = mu.compute_fine_Rt_fused(f1, f2, p1, p2, model, 0.1, 0.15)
    else:
        R, t, s = mu.compute_fine_Rt(mu.compute_feature_similarity(f1, f2, "cosine", 0.1, True), p1, p2, model)
    assert oracle.rotation_angle_deg(R.cpu(), torch.from_numpy(g["R"])).max().item() < 0.1   # <= 0.1 deg
    assert (t.cpu() - torch.from_numpy(g["t"])).abs().max().item() < 1e-4                     # <= 0.1 mm @ 1 m
    assert (s.cpu() - torch.from_numpy(g["score"])).abs().max().item() < 5e-3


def test_procrustes_golden(pkg, oracle, dev):
    g = _g("procrustes.npz")
    R, t = pkg.model_utils.weighted_procrustes(_t(g["src"], dev), _t(g["ref"], dev), _t(g["w"], dev),
                                               float(g["weight_thresh"]))
    assert oracle.rotation_angle_deg(R.cpu(), torch.from_numpy(g["R"])).max().item() < 0.01
    assert (t.cpu() - torch.from_numpy(g["t"])).abs().max().item() < 1e-5
