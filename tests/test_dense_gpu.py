"""GPU parity of the dense geometric stage (similarity, coarse and fine pose) through the
C ABI vs the CPU oracle, following the staged plan of SURVEY.md section 8c:

  S1  atten -> masks / cdf                      masks equal, cdf abs 2e-6 (values in [0,1])
  S2  oracle cdf + rand -> correspondences      bit exact
  S3  same triplets -> per-hypothesis R,t       1e-4 abs on well-conditioned triplets
  S4  same hypotheses -> scores                 rel 1e-4
  S5  end to end                                R <= 0.1 deg, t <= 1e-4 (radius units; 0.1 mm
                                                 for a 1 m radius object), same proposal id
Floating-point tolerances are stated next to each assert.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _coarse_inputs(pkg, oracle, ids, npoint=196, n_dense=2048, feat_dim=256, P1=6000, n_model=1024):
    b = pkg.synth.make_batch(ids, n_dense=n_dense, n_model=n_model, feat_dim=feat_dim, n_proposal1=P1)
    sp, sf, _ = oracle.sample_pts_feats(b["dense_pm"].numpy(), b["dense_fm"].numpy(), npoint)
    tp, tf, _ = oracle.sample_pts_feats(b["dense_po"].numpy(), b["dense_fo"].numpy(), npoint)
    sp, sf, tp, tf = (torch.from_numpy(np.ascontiguousarray(x)) for x in (sp, sf, tp, tf))
    bg = b["bg"][:, None, :]
    atten = oracle.feature_similarity(torch.cat([bg, sf], 1), torch.cat([bg, tf], 1), 0.1)
    return b, sp, tp, atten


@pytest.mark.parametrize("B,N1,N2,C", [(2, 197, 197, 256), (1, 2049, 2049, 256), (3, 50, 70, 64), (1, 130, 513, 32)])
def test_similarity(pkg, oracle, dev, B, N1, N2, C):
    g = torch.Generator().manual_seed(N1)
    f1, f2 = torch.randn(B, N1, C, generator=g), torch.randn(B, N2, C, generator=g)
    f1[0, 0] = 0.0  # zero row: F.normalize's eps path
    want = oracle.feature_similarity(f1, f2, 0.1)
    got = pkg.model_utils.compute_feature_similarity(f1.to(dev), f2.to(dev), "cosine", 0.1, True).cpu()
    # logits are in [-10, 10]; fp32 accumulation-order noise only
    assert (got - want).abs().max().item() < 3e-5
    raw = pkg.model_utils.compute_feature_similarity(f1.to(dev), f2.to(dev), "cosine", 1.0, False).cpu()
    ref = torch.bmm(f1, f2.transpose(1, 2))
    assert (raw - ref).abs().max().item() < 1e-4 * ref.abs().max().item()


def test_coarse_assign_S1(pkg, oracle, dev):
    b, sp, tp, atten = _coarse_inputs(pkg, oracle, [10, 11])
    want = oracle.coarse_assignment(atten)
    _, _, st = pkg.model_utils.compute_coarse_Rt(atten.to(dev), sp.to(dev), tp.to(dev), b["model"].to(dev),
                                                 6000, 300, rand=b["rand"].to(dev), return_stages=True)
    assert torch.equal(st["w1"].cpu(), want["w1"]) and torch.equal(st["w2"].cpu(), want["w2"])
    assert 0 < want["w1"].mean() < 1  # the masks are exercised
    # cdf values are in [0,1]: exp/pow rounding differences only
    assert (st["cdf"].cpu() - want["cdf"]).abs().max().item() < 2e-6
    assert (st["cdf"].cpu()[:, 1:] >= st["cdf"].cpu()[:, :-1]).all()


def test_coarse_hypotheses_S2_S3(pkg, oracle, dev):
    b, sp, tp, atten = _coarse_inputs(pkg, oracle, [12])
    a = oracle.coarse_assignment(atten)
    want = oracle.coarse_hypotheses(a["cdf"], b["rand"], sp, tp, 6000)
    L, ext = pkg._lib.lib(), pkg.pointnet2_ext
    B, P1, N = 1, 6000, 196
    cdf, rand = a["cdf"].contiguous().to(dev), b["rand"].to(dev)
    p1, p2 = sp.to(dev), tp.to(dev)
    R = torch.empty(B, P1, 3, 3, device=dev)
    t = torch.empty(B, P1, 3, device=dev)
    resid = torch.empty(B, P1, device=dev)
    corr = torch.empty(B, P1 * 3, dtype=torch.int32, device=dev)
    pkg._lib.check(L.pem_coarse_hypotheses(cdf.data_ptr(), rand.data_ptr(), p1.data_ptr(), p2.data_ptr(), B, N, N,
                                           P1, R.data_ptr(), t.data_ptr(), resid.data_ptr(), corr.data_ptr(),
                                           ext._stream()), "hyp")
    # S2: identical cdf + rand -> identical correspondences, bit exact
    assert torch.equal(corr.cpu().long(), want["idx"])
    # S3: well-conditioned triplets (sigma2/sigma1 > 1e-2) agree to 1e-4 absolute
    S = torch.linalg.svdvals(want["H"].double())
    ok = (S[..., 1] / S[..., 0].clamp_min(1e-30)) > 1e-2
    assert ok.float().mean() > 0.5
    dR = (R.cpu() - want["R"]).abs().flatten(2).max(2)[0]
    dt = (t.cpu() - want["t"]).abs().max(2)[0]
    assert dR[ok].max().item() < 1e-4 and dt[ok].max().item() < 1e-4
    assert (resid.cpu() - want["resid"])[ok].abs().max().item() < 1e-5
    # every returned R is a proper rotation, conditioned or not
    Rg = R.cpu().double()
    eye = torch.eye(3, dtype=torch.float64)
    assert ((Rg @ Rg.transpose(-1, -2)) - eye).abs().max().item() < 1e-5
    assert (torch.det(Rg) - 1).abs().max().item() < 1e-5


@pytest.mark.parametrize("N1,N2", [(20, 20), (196, 196), (300, 300)])
def test_sampler_edge_cases_on_device(pkg, oracle, dev, N1, N2):
    """weighted_sampling_onnx_compatible (MU:277-305) = argmax(cumsum >= rand): the FIRST element >= the draw,
    and 0 when none qualifies (argmax of an all-false row).  Cases: draws above the last cdf value (the cdf
    ends at S/(S+1e-8) < 1), draws equal to a cdf entry, runs of equal entries, a draw of 0, and an
    all-zero cdf (no foreground pair at all).  E = 90000 exceeds the kernel's shared-memory index
    (64 x 1024 entries), so the direct search of the tail is covered too.  Bit exact against the
    reference's compare/argmax form (faithful=True) and the lower-bound form."""
    g = torch.Generator().manual_seed(N1)
    E, P1 = N1 * N2, 512
    inc = torch.rand(3, E, generator=g)
    inc[:, torch.rand(E, generator=g) < 0.3] = 0.0     # runs of equal cdf entries (masked-out pairs)
    cdf = torch.cumsum(inc.double(), 1)
    cdf = (cdf / cdf[:, -1:] * 0.9).float()            # ends at 0.9: draws above it find nothing
    cdf[2] = 0.0                                       # no foreground pair: everything maps to 0
    rand = torch.rand(3, 3 * P1, generator=g)
    rand[:, 0] = 1.0
    rand[:, 1] = 0.95
    rand[:, 2] = 0.0
    rand[:, 3] = 0.9
    pick = torch.randint(0, E, (3, 200), generator=g)
    rand[:, 4:204] = torch.gather(cdf, 1, pick)        # draws that hit a cdf entry exactly
    rand[:, 204:300] = torch.nextafter(rand[:, 4:100], torch.tensor(2.0))
    want = oracle.inverse_cdf(cdf, rand, faithful=True)
    assert torch.equal(want, oracle.inverse_cdf(cdf, rand, faithful=False))
    assert (want[:, :2] == 0).all() and (want[2, 3:] == 0).all() and (want[0, 4:204] <= pick[0]).all()
    L, ext = pkg._lib.lib(), pkg.pointnet2_ext
    p1 = torch.randn(3, N1, 3, generator=g).to(dev)
    p2 = torch.randn(3, N2, 3, generator=g).to(dev)
    R, t = torch.empty(3, P1, 3, 3, device=dev), torch.empty(3, P1, 3, device=dev)
    resid = torch.empty(3, P1, device=dev)
    corr = torch.empty(3, P1 * 3, dtype=torch.int32, device=dev)
    cd, rd = cdf.to(dev), rand.to(dev)
    pkg._lib.check(L.pem_coarse_hypotheses(cd.data_ptr(), rd.data_ptr(), p1.data_ptr(), p2.data_ptr(), 3, N1, N2,
                                           P1, R.data_ptr(), t.data_ptr(), resid.data_ptr(), corr.data_ptr(),
                                           ext._stream()), "hyp")
    assert torch.equal(corr.cpu().long(), want)
    # instance 2 draws correspondence (0,0) three times.  H is NOT zero there: the reference's weights are
    # 1/(3+1e-5) (MU:380-384), so the "centroid" is p*3/(3+1e-5) and the centred points are three copies of
    # p*1e-5/3 -> a rank-1 H of size 1e-11 whose rotation is anybody's (degenerate triplets).  What is
    # defined: a proper rotation that maps the correspondence onto itself (the exact H = 0 -> R = I convention
    # is asserted on the operator seam, tests/test_operator_seam_gpu.py).
    Rg = R[2].cpu().double()
    assert ((Rg @ Rg.transpose(-1, -2)) - torch.eye(3, dtype=torch.float64)).abs().max().item() < 1e-5
    assert (torch.det(Rg) - 1).abs().max().item() < 1e-5
    assert resid[2].abs().max().item() < 1e-4
    assert torch.isfinite(R).all() and torch.isfinite(t).all() and torch.isfinite(resid).all()
    # the standalone operator (CustomSearchSorted seam) agrees where an element qualifies
    ss = pkg.model_utils.searchsorted(cd, rd).cpu()
    assert torch.equal(ss, torch.searchsorted(cdf, rand))
    assert torch.equal(torch.where(ss >= E, torch.zeros_like(ss), ss), want)


@pytest.mark.parametrize("P1,P2", [(6000, 300), (1000, 1000), (777, 1), (60000, 300)])
def test_select_smallest(pkg, dev, P1, P2):
    g = torch.Generator().manual_seed(P1)
    v = torch.rand(3, P1, generator=g)
    v[0, :50] = 0.0           # ties at the bottom (degenerate triplets have residual exactly 0)
    v[1, 5] = float("nan")    # NaN sorts last
    v[2] = torch.round(v[2] * 20) / 20  # heavy ties everywhere
    sel = torch.empty(3, P2, dtype=torch.int32, device=dev)
    vd = v.to(dev)
    pkg._lib.check(pkg._lib.lib().pem_select_smallest(vd.data_ptr(), 3, P1, P2, sel.data_ptr(),
                                                      pkg.pointnet2_ext._stream()), "select")
    sel = sel.cpu().long()
    for b in range(3):
        s = sel[b]
        assert len(set(s.tolist())) == P2 and (s[1:] > s[:-1]).all()  # a set, ascending ids
        vv = torch.nan_to_num(v[b], nan=float("inf"))
        kth = torch.sort(vv)[0][P2 - 1]
        assert (vv[s] <= kth).all()                      # nothing above the threshold
        assert (vv < kth).sum() == (vv[s] < kth).sum()   # everything below it is in
        ties = torch.nonzero(vv == kth).flatten()
        chosen = s[vv[s] == kth]
        assert torch.equal(chosen, ties[: len(chosen)])  # lowest ids among the ties


def test_score_hypotheses_S4(pkg, oracle, dev):
    b, sp, tp, atten = _coarse_inputs(pkg, oracle, [13, 14])
    a = oracle.coarse_assignment(atten)
    h = oracle.coarse_hypotheses(a["cdf"], b["rand"], sp, tp, 6000)
    R, t = h["R"][:, :500].contiguous(), h["t"][:, :500].contiguous()
    want = oracle.score_hypotheses(R, t, sp, a["w1"], b["model"])
    got = pkg.model_utils.score_hypotheses(R.to(dev), t.to(dev), sp.to(dev), a["w1"].to(dev), b["model"].to(dev)).cpu()
    assert ((got - want).abs() / want.abs().clamp_min(1e-6)).max().item() < 1e-4
    # with a selection list and 2048 model points (the sweep configuration)
    model2 = pkg.synth.make_batch([13, 14], n_model=2048)["model"]
    sel = torch.randperm(500)[:77].repeat(2, 1).to(torch.int32)
    want2 = oracle.score_hypotheses(R[:, sel[0].long()], t[:, sel[0].long()], sp, a["w1"], model2)
    got2 = pkg.model_utils.score_hypotheses(R.to(dev), t.to(dev), sp.to(dev), a["w1"].to(dev), model2.to(dev),
                                            sel=sel.to(dev)).cpu()
    assert ((got2 - want2).abs() / want2.abs().clamp_min(1e-6)).max().item() < 1e-4


@pytest.mark.parametrize("Nm,H,n1,nfg", [(1024, 300, 196, 150), (1000, 37, 196, 196), (300, 5, 50, 1),
                                           (1024, 1, 200, 0), (513, 410, 256, 131), (2048, 64, 196, 100),
                                           (128, 600, 196, 150), (100, 900, 64, 64), (3000, 50, 196, 120),
                                           (1024, 100, 256, 256), (257, 333, 64, 33)])
def test_score_hypotheses_tensor_core_vs_simt_and_oracle(pkg, oracle, dev, Nm, H, n1, nfg):
    """The tcgen05 (split-tf32) scoring kernel against the SIMT kernel and the oracle: ragged model
    sizes, hypothesis counts that do not divide over the CTAs, several batches per CTA, no / one /
    all foreground points, one accumulator chunk per tile (Nm <= 128: the builders run furthest ahead of the
    reducers), several model blocks (Nm > 1024), 256 foreground points (smaller batches)."""
    import ctypes
    g = torch.Generator().manual_seed(Nm + H)
    Bn = 3
    q, _ = torch.linalg.qr(torch.randn(Bn, H, 3, 3, generator=g))
    R = q.contiguous()
    t = (0.1 * torch.randn(Bn, H, 3, generator=g)).contiguous()
    pts = (0.4 * torch.randn(Bn, n1, 3, generator=g)).contiguous()
    model = (0.4 * torch.randn(Bn, Nm, 3, generator=g)).contiguous()
    w1 = torch.zeros(Bn, n1)
    for b in range(Bn):
        w1[b, torch.randperm(n1, generator=g)[:nfg]] = 1.0
    want = oracle.score_hypotheses(R, t, pts, w1, model)
    args = (R.to(dev), t.to(dev), pts.to(dev), w1.to(dev), model.to(dev))
    simt = pkg.model_utils.score_hypotheses(*args, kernel="simt").cpu()
    tc = pkg.model_utils.score_hypotheses(*args, kernel="tensor").cpu()
    assert torch.equal(tc, pkg.model_utils.score_hypotheses(*args).cpu())   # deterministic sums
    if nfg == 0:
        assert (tc == 0).all() and (simt == 0).all()
        return
    rel = lambda a, b: ((a - b).abs() / b.abs().clamp_min(1e-6)).max().item()
    assert rel(simt, want) < 1e-4
    assert rel(tc, want) < 1e-4
    assert rel(tc, simt) < 2e-5
    assert torch.equal(tc.argmax(1), simt.argmax(1)) or rel(tc.max(1)[0], simt.max(1)[0]) < 1e-6


def test_coarse_end_to_end_S5(pkg, oracle, dev):
    """From atten to the coarse pose.

    What "identical selected hypothesis" can mean here is limited by the reference itself:
    about a quarter of its 6000 triplets draw the same correspondence twice, their H is rank 1
    up to fp32 rounding noise (sigma2/sigma1 ~ 1e-8), and the rotation about the remaining axis
    that LAPACK returns for them is a function of that noise (an exact SVD of the reference's own
    fp32 H differs from its R by > 0.3 rad; DESIGN.md, "degenerate triplets").  Such hypotheses
    have tiny residuals, fill much of the top-300 and often win.  So the check is:
      * the top-P2 sets agree (up to residual ties at the threshold),
      * on the well-posed hypotheses (sigma2/sigma1 > 1e-3) of the common set, R, t and the score
        agree and the best one is the same proposal id,
      * given the REFERENCE's hypotheses, selection + scoring + argmax return the identical winner
        bit for bit (test_coarse_selection_and_argmax_given_reference_hypotheses)."""
    ids = [20, 21, 22, 23]
    b, sp, tp, atten = _coarse_inputs(pkg, oracle, ids)
    check_coarse_against_oracle(pkg, oracle, dev, atten, sp, tp, b["model"], 6000, 300, b["rand"])


def check_coarse_against_oracle(pkg, oracle, dev, atten, sp, tp, model, P1, P2, rand, min_common=20, stats=None):
    """Staged comparison of compute_coarse_Rt with the oracle on the same atten / points / draw.

    Bars: S3 on hypotheses whose H has sigma2/sigma1 > 1e-2 and whose three correspondences were
    drawn identically: R, t 1e-4 abs.  S4: OUR scores against the oracle's scorer applied to OUR hypotheses
    (identical hypotheses): rel 1e-4; end to end against the oracle's own score of the same proposal id
    (hypotheses differ by the S3 error, |dR| ~ 1e-5 against point-to-model distances of ~ 0.03): rel 1e-3.
    The best well-posed hypothesis is the same proposal id unless the oracle's top-2 gap is below 1e-3.
    `stats` (dict) collects the winner-id match counts for the report under profiles/."""
    atten, sp, tp, model, rand = (torch.as_tensor(x) for x in (atten, sp, tp, model, rand))
    n_inst = sp.shape[0]
    oR, ot, ost = oracle.coarse_Rt(atten, sp, tp, model, P1, P2, rand)
    R, t, st = pkg.model_utils.compute_coarse_Rt(atten.to(dev), sp.to(dev), tp.to(dev), model.to(dev), P1,
                                                 P2, rand=rand.to(dev), return_stages=True)
    st = {k: v.cpu() for k, v in st.items()}
    S = torch.linalg.svdvals(ost["H"].double())
    ratio = S[..., 1] / S[..., 0].clamp_min(1e-30)
    posed = ratio > 1e-2
    worst = {"dR": 0.0, "dt": 0.0, "score_staged": 0.0, "score_e2e": 0.0}
    n_same_winner = n_winner_posed = n_same_posed_winner = 0
    for i in range(n_inst):
        ours, theirs = st["sel"][i].long(), ost["top"][i]
        assert len(set(ours.tolist()) ^ set(theirs.tolist())) <= 8
        o_score = dict(zip(theirs.tolist(), ost["scores"][i].tolist()))
        g_score = dict(zip(ours.tolist(), st["scores"][i].tolist()))
        same_draw = (st["corr"][i].long() == ost["idx"][i]).view(-1, 3).all(1)
        common = [h for h in ours.tolist() if h in o_score and posed[i, h] and same_draw[h]]
        assert len(common) >= min_common
        ci = torch.tensor(common)
        # S4 staged: the oracle's scorer on OUR hypotheses
        staged = oracle.score_hypotheses(st["R"][i, ci].view(1, -1, 3, 3), st["t"][i, ci].view(1, -1, 3), sp[i:i + 1],
                                         ost["w1"][i:i + 1], model[i:i + 1])[0]
        for k, h in enumerate(common):
            worst["dR"] = max(worst["dR"], (st["R"][i, h].view(3, 3) - ost["R"][i, h]).abs().max().item())
            worst["dt"] = max(worst["dt"], (st["t"][i, h] - ost["t"][i, h]).abs().max().item())
            worst["score_staged"] = max(worst["score_staged"], abs(g_score[h] - staged[k].item()) / staged[k].item())
            worst["score_e2e"] = max(worst["score_e2e"], abs(g_score[h] - o_score[h]) / o_score[h])
        best_o = max(common, key=lambda h: o_score[h])
        best_g = max(common, key=lambda h: g_score[h])
        second = sorted((o_score[h] for h in common), reverse=True)[1]
        assert best_o == best_g or (o_score[best_o] - second) / o_score[best_o] < 1e-3, (i, best_o, best_g)
        same = st["proposal_id"][i].item() == ost["proposal_id"][i].item()
        n_same_winner += int(same)
        w_posed = bool(posed[i, ost["proposal_id"][i]])
        n_winner_posed += int(w_posed)
        n_same_posed_winner += int(same and w_posed)
        if same and w_posed:  # identical well-posed winner -> pose within tolerance
            assert oracle.rotation_angle_deg(R[i].cpu(), oR[i]).item() < 0.1
            assert (t[i].cpu() - ot[i]).abs().max().item() < 1e-4
        # (no assertion on the overall best score: when a degenerate triplet wins, its score is a
        #  function of LAPACK's rounding noise on one side and of ours on the other)
    print("coarse winner identical on %d of %d instances; the oracle's winner is well-posed on %d, identical on %d "
          "of those; worst well-posed errors %s" % (n_same_winner, n_inst, n_winner_posed, n_same_posed_winner, worst))
    if stats is not None:
        stats["instances"] = stats.get("instances", 0) + n_inst
        stats["winner_id_match"] = stats.get("winner_id_match", 0) + n_same_winner
        stats["oracle_winner_well_posed"] = stats.get("oracle_winner_well_posed", 0) + n_winner_posed
        stats["winner_id_match_given_well_posed"] = stats.get("winner_id_match_given_well_posed", 0) + n_same_posed_winner
        for k, v in worst.items():
            stats["worst_" + k] = max(stats.get("worst_" + k, 0.0), v)
    assert worst["dR"] < 1e-4 and worst["dt"] < 1e-4, worst          # S3
    assert worst["score_staged"] < 1e-4, worst                          # S4, identical hypotheses
    assert worst["score_e2e"] < 1e-3, worst                             # end to end (S3 error propagated)
    # correspondences drawn from OUR cdf differ from the oracle's only for draws that land
    # within rounding distance of a bin edge
    assert (st["corr"].long() != ost["idx"]).float().mean().item() < 0.005
    return oR, ot


def test_coarse_selection_and_argmax_given_reference_hypotheses(pkg, oracle, dev):
    """S4+: with the REFERENCE's hypotheses and residuals as input, selection, scoring and the
    final argmax pick the identical proposal id."""
    ids = [24, 25, 26]
    b, sp, tp, atten = _coarse_inputs(pkg, oracle, ids)
    _, _, ost = oracle.coarse_Rt(atten, sp, tp, b["model"], 6000, 300, b["rand"])
    L, stream = pkg._lib.lib(), pkg.pointnet2_ext._stream()
    B, P1, P2 = len(ids), 6000, 300
    Rh, th = ost["R"].contiguous().to(dev), ost["t"].contiguous().to(dev)
    resid = ost["resid"].contiguous().to(dev)
    sel = torch.empty(B, P2, dtype=torch.int32, device=dev)
    pkg._lib.check(L.pem_select_smallest(resid.data_ptr(), B, P1, P2, sel.data_ptr(), stream), "select")
    for i in range(B):  # identical sets, except that ties AT the threshold value may be broken differently
        ours, theirs = set(sel[i].cpu().tolist()), set(ost["top"][i].tolist())
        kth = torch.sort(ost["resid"][i])[0][P2 - 1]
        assert all(ost["resid"][i, h] == kth for h in ours ^ theirs), (i, len(ours ^ theirs))
    scores = pkg.model_utils.score_hypotheses(Rh, th, sp.to(dev), ost["w1"].to(dev), b["model"].to(dev), sel=sel)
    Ro, to = torch.empty(B, 9, device=dev), torch.empty(B, 3, device=dev)
    pid = torch.empty(B, dtype=torch.int32, device=dev)
    pkg._lib.check(L.pem_select_best(scores.data_ptr(), sel.data_ptr(), Rh.data_ptr(), th.data_ptr(), B, P1, P2,
                                     Ro.data_ptr(), to.data_ptr(), pid.data_ptr(), stream), "best")
    for i in range(B):
        top2 = torch.topk(ost["scores"][i], 2)[0]
        tie = (top2[0] - top2[1]) / top2[0] < 1e-4
        assert pid[i].item() == ost["proposal_id"][i].item() or tie
        if pid[i].item() == ost["proposal_id"][i].item():  # identical winner -> identical pose, bit for bit
            assert torch.equal(Ro[i].cpu().view(3, 3), ost["R"][i, pid[i].item()])
            assert torch.equal(to[i].cpu(), ost["t"][i, pid[i].item()])


def _fine_case(pkg, oracle, ids, n_dense, feat_dim, n_model=1024):
    b = pkg.synth.make_batch(ids, n_dense=n_dense, n_model=n_model, feat_dim=feat_dim, n_proposal1=16)
    bg = b["bg"][:, None, :]
    atten = oracle.feature_similarity(torch.cat([bg, b["dense_fm"]], 1), torch.cat([bg, b["dense_fo"]], 1), 0.1)
    return b, atten


@pytest.mark.parametrize("ids,n_dense,feat_dim", [([30, 31], 512, 64), ([32], 2048, 256), ([33, 34, 35], 200, 32)])
def test_fine_Rt(pkg, oracle, dev, ids, n_dense, feat_dim):
    b, atten = _fine_case(pkg, oracle, ids, n_dense, feat_dim)
    oR, ot, os_, ost = oracle.fine_Rt(atten, b["dense_pm"], b["dense_po"], b["model"])
    R, t, s, st = pkg.model_utils.compute_fine_Rt(atten.to(dev), b["dense_pm"].to(dev), b["dense_po"].to(dev),
                                                  b["model"].to(dev), return_stages=True)
    # masks: identical except for points whose foreground/background decision is a near tie
    for k in ("w1", "w2"):
        assert (st[k].cpu() != ost[k]).float().mean().item() < 2e-3
    assert 0 < ost["w1"].mean() < 1
    assert (st["row"].cpu() - ost["row"]).abs().max().item() < 1e-4       # row sums are in [0,1]
    assert oracle.rotation_angle_deg(R.cpu(), oR).max().item() < 0.1     # north_star: <= 0.1 deg
    assert (t.cpu() - ot).abs().max().item() < 1e-4                      # <= 0.1 mm at 1 m radius
    assert (s.cpu() - os_).abs().max().item() < 2e-3                     # one point of 2048 flips 5e-4
    # and the planted pose is recovered
    if n_dense >= 512:
        assert oracle.rotation_angle_deg(R.cpu(), b["gt_R"]).max().item() < 1.0


def test_fine_Rt_generic_logits(pkg, oracle, dev):
    """atten need not come from cosine similarity: large-range logits must not overflow."""
    g = torch.Generator().manual_seed(3)
    atten = torch.randn(2, 301, 257, generator=g) * 30.0
    p1, p2 = torch.randn(2, 300, 3, generator=g) * 0.3, torch.randn(2, 256, 3, generator=g) * 0.3
    oR, ot, os_, ost = oracle.fine_Rt(atten, p1, p2, None)
    R, t, s, st = pkg.model_utils.compute_fine_Rt(atten.to(dev), p1.to(dev), p2.to(dev), None, return_stages=True)
    assert torch.isfinite(R).all() and torch.isfinite(t).all()
    assert (st["w1"].cpu() != ost["w1"]).float().mean().item() < 5e-3
    assert (st["row"].cpu() - ost["row"]).abs().max().item() < 1e-4
    assert oracle.rotation_angle_deg(R.cpu(), oR).max().item() < 0.1
    assert (t.cpu() - ot).abs().max().item() < 1e-4


def test_weighted_procrustes(pkg, oracle, dev):
    g = torch.Generator().manual_seed(9)
    src = torch.randn(5, 400, 3, generator=g)
    Rg = torch.linalg.qr(torch.randn(5, 3, 3, generator=g))[0]
    Rg = Rg * torch.sign(torch.det(Rg)).view(-1, 1, 1)
    ref = src @ Rg.transpose(1, 2) + torch.randn(5, 1, 3, generator=g) + 0.01 * torch.randn(5, 400, 3, generator=g)
    w = torch.rand(5, 400, generator=g)
    for thresh, ww in ((0.0, w), (0.5, w), (0.5, None)):
        oR, ot, _ = oracle.procrustes(src, ref, ww, thresh)
        R, t = pkg.model_utils.weighted_procrustes(src.to(dev), ref.to(dev), None if ww is None else ww.to(dev), thresh)
        assert oracle.rotation_angle_deg(R.cpu(), oR).max().item() < 0.01
        assert (t.cpu() - ot).abs().max().item() < 1e-5
    # reflection case: det-correction must yield a proper rotation
    refl = src.clone()
    refl[..., 0] *= -1
    oR, ot, _ = oracle.procrustes(src, refl, None, 0.0)
    R, t = pkg.model_utils.weighted_procrustes(src.to(dev), refl.to(dev), None, 0.0)
    assert (torch.det(R.cpu()) - 1).abs().max().item() < 1e-5
    assert oracle.rotation_angle_deg(R.cpu(), oR).max().item() < 0.05
