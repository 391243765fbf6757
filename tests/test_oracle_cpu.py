"""CPU: pin the oracle (oracle/) against (a) the committed golden vectors written from the live
reference by oracle/make_golden.py and (b) the live reference itself when /root/reference is
present (this container; skipped on the GPU box).  Integer/byte work bit exact; the dense
functions are torch-CPU restatements of torch-CPU reference code and are bit exact too."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT


def _g(name):
    return np.load(os.path.join(GOLDEN, name))


def test_pointnet2_small_golden(oracle):
    g = _g("pointnet2_small.npz")
    np.testing.assert_array_equal(oracle.fps(g["xyz"], 96), g["fps_idx"])
    assert 11 not in g["fps_idx"]                                   # near-origin point never sampled
    np.testing.assert_array_equal(oracle.gather(g["feat"], g["gather_idx"]), g["gathered"])
    assert (g["gathered"][0, :, 0] == 0).all()                      # invalid index -> 0.0
    for key, r, ns in (("bq_r01_ns32", 0.1, 32), ("bq_r02_ns64", 0.2, 64)):
        got = oracle.ball_query(g["centres"], g["xyz"], r, ns)
        np.testing.assert_array_equal(got, g[key].astype(np.int32))
        assert (got[:, -1] == 0).all()                              # centre without neighbours
    np.testing.assert_array_equal(oracle.group(g["feat"], g["bq_r02_ns64"].astype(np.int32)), g["grouped"])
    qg, _ = oracle.query_and_group(g["xyz"], (torch.from_numpy(g["xyz"]) + 1e-8).numpy(),
                                   np.ascontiguousarray(g["xyz"].transpose(0, 2, 1)), 0.1, 32)
    np.testing.assert_array_equal(qg, g["qg_r01_ns32"])


def test_pointnet2_default_golden(oracle):
    g = _g("pointnet2_default.npz")
    np.testing.assert_array_equal(oracle.fps(g["dense_pm"], 196), g["fps_m"].astype(np.int32))
    np.testing.assert_array_equal(oracle.fps(g["dense_po"], 196), g["fps_o"].astype(np.int32))
    po = torch.from_numpy(g["dense_po"])
    cen = (po + 1e-8).numpy()
    np.testing.assert_array_equal(oracle.ball_query(cen, g["dense_po"], 0.1, 32), g["bq_r01_ns32"].astype(np.int32))
    np.testing.assert_array_equal(oracle.ball_query(cen, g["dense_po"], 0.2, 64), g["bq_r02_ns64"].astype(np.int32))


@pytest.mark.parametrize("name", ["coarse_small.npz", "coarse_default.npz"])
def test_coarse_golden(oracle, name):
    g = _g(name)
    p1n = int(g["n_proposal1"]) if "n_proposal1" in g else 6000
    p2n = int(g["n_proposal2"]) if "n_proposal2" in g else 300
    if "f1" in g:
        atten = oracle.feature_similarity(g["f1"], g["f2"], 0.1)
        assert torch.equal(atten, torch.from_numpy(g["atten"]))
    R, t, st = oracle.coarse_Rt(g["atten"], g["pts1"], g["pts2"], g["model"], p1n, p2n, g["rand"])
    assert torch.equal(R, torch.from_numpy(g["R"])) and torch.equal(t, torch.from_numpy(g["t"]))
    # the lower-bound sampler equals the reference's compare/argmax sampler
    idx_f = oracle.inverse_cdf(st["cdf"], g["rand"], faithful=True)
    assert torch.equal(idx_f, st["idx"])


def test_fine_golden(oracle):
    g = _g("fine_small.npz")
    atten = oracle.feature_similarity(g["f1"], g["f2"], 0.1)
    R, t, s, _ = oracle.fine_Rt(atten, g["pts1"], g["pts2"], g["model"])
    assert torch.equal(R, torch.from_numpy(g["R"])) and torch.equal(t, torch.from_numpy(g["t"]))
    assert torch.equal(s, torch.from_numpy(g["score"]))


def test_procrustes_golden(oracle):
    g = _g("procrustes.npz")
    R, t, _ = oracle.procrustes(g["src"], g["ref"], g["w"], float(g["weight_thresh"]))
    assert torch.equal(R, torch.from_numpy(g["R"])) and torch.equal(t, torch.from_numpy(g["t"]))


def test_cumsum_is_fp64_accumulated(oracle):
    """The device scan assumes torch's CPU cumsum keeps a double accumulator."""
    x = torch.rand(3, 38416) ** 8
    assert torch.equal(torch.cumsum(x, 1), torch.cumsum(x.double(), 1).float())


def test_sampler_returns_zero_when_nothing_qualifies(oracle):
    cdf = torch.tensor([[0.1, 0.4, 0.4, 0.9]])
    r = torch.tensor([[0.05, 0.1, 0.4, 0.41, 0.95, 0.0]])
    want = torch.tensor([[0, 0, 1, 3, 0, 0]])
    assert torch.equal(oracle.inverse_cdf(cdf, r), want)
    assert torch.equal(oracle.inverse_cdf(cdf, r, faithful=True), want)


def test_rotation_angle_helper(oracle):
    a = torch.eye(3)
    th = np.deg2rad(0.05)
    b = torch.tensor([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    assert abs(oracle.rotation_angle_deg(a, b).item() - 0.05) < 1e-6


needs_ref = pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="live reference not present")


@needs_ref
def test_oracle_matches_live_reference(oracle):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_live
    mu, pu, ext = ref_live.load()
    g = torch.Generator().manual_seed(5)
    pts = torch.randn(2, 1500, 3, generator=g) * 0.3
    pts[:, 3] = 0.002
    pts[:, 700:720] = pts[:, 20:40]
    np.testing.assert_array_equal(ext.furthest_point_sampling(pts, 200).numpy(), oracle.fps(pts.numpy(), 200))
    np.testing.assert_array_equal(ext.ball_query(pts + 1e-8, pts, 0.2, 64).numpy(),
                                  oracle.ball_query((pts + 1e-8).numpy(), pts.numpy(), 0.2, 64))
    f1, f2 = torch.randn(2, 197, 64, generator=g), torch.randn(2, 197, 64, generator=g)
    atten = mu.compute_feature_similarity(f1, f2, "cosine", 0.1, True)
    assert torch.equal(atten, oracle.feature_similarity(f1, f2, 0.1))
    p1, p2 = torch.randn(2, 196, 3, generator=g) * 0.3, torch.randn(2, 196, 3, generator=g) * 0.3
    model = torch.randn(2, 512, 3, generator=g) * 0.3
    torch.manual_seed(3)
    state = torch.get_rng_state()
    R, t = mu.compute_coarse_Rt(atten, p1, p2, model, 2000, 100)
    torch.set_rng_state(state)
    Ro, to, _ = oracle.coarse_Rt(atten, p1, p2, model, 2000, 100, torch.rand(2, 6000))
    assert torch.equal(R, Ro) and torch.equal(t, to)
    af = torch.randn(2, 301, 301, generator=g) * 4
    q1, q2 = torch.randn(2, 300, 3, generator=g) * 0.3, torch.randn(2, 300, 3, generator=g) * 0.3
    Rf, tf, sf = mu.compute_fine_Rt(af, q1, q2, model)
    Rg, tg, sg, _ = oracle.fine_Rt(af, q1, q2, model)
    assert torch.equal(Rf, Rg) and torch.equal(tf, tg) and torch.equal(sf, sg)


@needs_ref
def test_cpu_stage_uses_the_reference_extension(oracle, pkg):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_stage
    ops = cpu_stage.pointnet2_ops()
    assert ops.kind == "reference"
    b = pkg.synth.make_batch([0], n_dense=256, n_model=128, feat_dim=32, n_proposal1=200)
    R, t, s, st = cpu_stage.run_stage_cpu(b, n_sparse=32, n_proposal1=200, n_proposal2=20, faithful=True, ops=ops)
    R2, t2, s2, _ = cpu_stage.run_stage_cpu(b, n_sparse=32, n_proposal1=200, n_proposal2=20, faithful=False,
                                            ops=cpu_stage._PortOps())
    assert torch.equal(R, R2) and torch.equal(t, t2) and torch.equal(s, s2)
