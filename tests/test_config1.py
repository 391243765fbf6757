"""BASELINE.json configs[0]: the bundled Data/Example instance (obj_000005 CAD + depth image)
pushed through the REFERENCE's own CoarsePointMatching / FinePointMatching modules with
random-init weights on CPU (oracle/make_config1.py); every tensor crossing the geometric-function
boundary was recorded into tests/golden/config1_example.npz.  CPU: the oracle reproduces the
recorded outputs bit for bit.  GPU: the CUDA path reproduces them within the stated tolerances."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN


@pytest.fixture(scope="module")
def c1():
    g = np.load(os.path.join(GOLDEN, "config1_example.npz"))
    d = {k: torch.from_numpy(np.ascontiguousarray(g[k])) for k in g.files}
    torch.manual_seed(1)
    d["dense_fm"], d["dense_fo"] = torch.randn(1, 2048, 256), torch.randn(1, 2048, 256)
    chk = torch.stack([d["dense_fm"].double().sum(), d["dense_fo"].double().sum(),
                       d["dense_fm"][0, 7, 3].double(), d["dense_fo"][0, 2047, 255].double()])
    assert torch.equal(chk, d["dense_f_checksum"]), "torch RNG stream changed: regenerate the fixture"
    return d


def test_oracle_reproduces_reference_on_config1(oracle, c1):
    sp, sf, idx = oracle.sample_pts_feats(c1["dense_pm"].numpy(), c1["dense_fm"].numpy(), 196)
    assert np.array_equal(idx, c1["fps_m"].numpy()) and np.array_equal(sp, c1["sparse_pm"].numpy())
    assert np.array_equal(sf, c1["sparse_fm"].numpy())
    atten = oracle.feature_similarity(c1["coarse_f1"], c1["coarse_f2"], 0.1)
    assert torch.equal(atten, c1["coarse_atten"])
    R0, t0, _ = oracle.coarse_Rt(c1["coarse_atten"], c1["sparse_pm"], c1["sparse_po"], c1["model_pts"], 6000, 300,
                                 c1["rand"])
    assert torch.equal(R0, c1["init_R"]) and torch.equal(t0, c1["init_t"])
    af = oracle.feature_similarity(c1["fine_f1"], c1["fine_f2"], 0.1)
    R, t, s, _ = oracle.fine_Rt(af, c1["dense_pm"], c1["dense_po"], c1["model_pts"])
    assert torch.equal(R, c1["fine_R"]) and torch.equal(t, c1["fine_t"]) and torch.equal(s, c1["fine_score"])


@pytest.mark.gpu
def test_cuda_path_on_config1(pkg, oracle, dev, c1):
    from test_dense_gpu import check_coarse_against_oracle
    mu, pu = pkg.model_utils, pkg.pointnet2_utils
    g = {k: v.to(dev) for k, v in c1.items() if v.dtype in (torch.float32, torch.int32)}
    # step 2: FPS + gathers on the real depth / CAD clouds: bit exact
    for pk, fk, ik, sk, sfk in (("dense_pm", "dense_fm", "fps_m", "sparse_pm", "sparse_fm"),
                                ("dense_po", "dense_fo", "fps_o", "sparse_po", "sparse_fo")):
        p, f, idx = mu.sample_pts_feats(g[pk], g[fk], 196, return_index=True)
        assert torch.equal(idx.cpu(), c1[ik]) and torch.equal(p.cpu(), c1[sk]) and torch.equal(f.cpu(), c1[sfk])
    # step 3: coarse similarity and pose
    att = mu.compute_feature_similarity(g["coarse_f1"], g["coarse_f2"], "cosine", 0.1, True)
    assert (att.cpu() - c1["coarse_atten"]).abs().max().item() < 3e-5
    oR, ot = check_coarse_against_oracle(pkg, oracle, dev, c1["coarse_atten"], c1["sparse_pm"], c1["sparse_po"],
                                         c1["model_pts"], 6000, 300, c1["rand"], min_common=3)
    assert torch.equal(oR, c1["init_R"])
    # step 4 front end: PE grouping of the transformed observed cloud: bit exact vs the CPU loops
    p1_ = ((c1["dense_pm"] - c1["init_t"].unsqueeze(1)) @ c1["init_R"]).contiguous()
    o1, o2 = pu.pe_group2(p1_.to(dev))
    w1, _ = oracle.query_and_group(p1_.numpy(), (p1_ + 1e-8).numpy(), p1_.transpose(1, 2).contiguous().numpy(), 0.1, 32)
    w2, _ = oracle.query_and_group(p1_.numpy(), (p1_ + 1e-8).numpy(), p1_.transpose(1, 2).contiguous().numpy(), 0.2, 64)
    assert np.array_equal(o1.cpu().numpy(), w1) and np.array_equal(o2.cpu().numpy(), w2)
    # step 4: fine pose, fused tensor-core path (BOTH arithmetics: split-bf16 and the stage's default, one fp16
    # product per element) and materialised path, vs the reference's output.  Random-init weights = the low-signal
    # regime (plain bf16 moved R by 0.046 deg here); tolerance as north_star states.
    w1_ref = oracle.fine_Rt(oracle.feature_similarity(c1["fine_f1"], c1["fine_f2"], 0.1), c1["dense_pm"],
                            c1["dense_po"], c1["model_pts"])[3]["w1"]
    report = {}
    for path in ("fused/split_bf16", "fused/fp16", "materialised"):
        if path.startswith("fused"):
            R, t, s, st = mu.compute_fine_Rt_fused(g["fine_f1"], g["fine_f2"], g["dense_pm"], g["dense_po"],
                                                   g["model_pts"], 0.1, 0.15, return_stages=True,
                                                   precision=path.split("/")[1])
        else:
            af = mu.compute_feature_similarity(g["fine_f1"], g["fine_f2"], "cosine", 0.1, True)
            R, t, s, st = mu.compute_fine_Rt(af, g["dense_pm"], g["dense_po"], g["model_pts"], return_stages=True)
        ang = oracle.rotation_angle_deg(R.cpu(), c1["fine_R"]).max().item()
        dt_mm = (t.cpu() - c1["fine_t"]).abs().max().item() * c1["radius"].item() * 1000.0
        flips = int((st["w1"].cpu() != w1_ref).sum().item())
        report[path] = {"dR_deg": ang, "dt_mm": dt_mm, "score": s.item(), "ref_score": c1["fine_score"].item(),
                        "foreground_label_flips_of_2048": flips}
        print("config1 %s: dR = %.5f deg, dt = %.5f mm, score %.4f vs %.4f, %d label flips"
              % (path, ang, dt_mm, s.item(), c1["fine_score"].item(), flips))
        assert ang < 0.1 and dt_mm < 0.1, (path, report)         # <= 0.1 deg, <= 0.1 mm
        assert abs(s.item() - c1["fine_score"].item()) < 5e-3, (path, report)
        assert flips <= 20, (path, report)                        # < 1 %: near-tie foreground decisions only (count reported)
    from conftest import write_report
    write_report("r2_config1_fine_parity.json", report)
