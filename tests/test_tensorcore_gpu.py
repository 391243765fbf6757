"""GPU parity of the tcgen05 (tensor-core) paths: the split-bf16 similarity GEMM and the
fused fine stage that never materialises the attention matrix, vs the CPU oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu

# Both GEMM arithmetics of the fused fine stage (include/pem_b200.h, PEM_PREC_*).  The FINAL pose bars
# are the north_star's and the same for both (R <= 0.1 deg, t <= 1e-4 radius units = 0.1 mm at 1 m,
# score 2e-3); only the bars on the intermediates follow the logit precision: split-bf16 logits are
# good to ~1e-5, one fp16 product to ~1e-3 of the +-10 range (assignment scores to ~2e-4 observed).
PRECISIONS = ["split_bf16", "fp16"]
ROW_TOL = {"split_bf16": 1e-4, "fp16": 1e-3}
PRED_TOL = {"split_bf16": 1e-3, "fp16": 2e-3}


@pytest.mark.parametrize("B,N1,N2,C", [(1, 129, 129, 64), (2, 2049, 2049, 256), (3, 197, 197, 128), (1, 130, 513, 192)])
def test_similarity_tensorcore(pkg, oracle, dev, B, N1, N2, C):
    g = torch.Generator().manual_seed(N1 + C)
    f1, f2 = torch.randn(B, N1, C, generator=g), torch.randn(B, N2, C, generator=g)
    f1[0, 0] = 0.0
    want = oracle.feature_similarity(f1, f2, 0.1)
    got = pkg.model_utils.compute_feature_similarity(f1.to(dev), f2.to(dev), "cosine", 0.1, True, backend="tc").cpu()
    # logits in [-10,10]; split-bf16 (hi*hi + lo*hi + hi*lo) keeps ~2^-17 relative per product
    assert (got - want).abs().max().item() < 1e-4


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("ids,n_dense,feat_dim", [([30, 31], 512, 64), ([32, 36, 37], 2048, 256), ([33, 34, 35], 200, 128)])
def test_fine_fused(pkg, oracle, dev, ids, n_dense, feat_dim, precision):
    b = pkg.synth.make_batch(ids, n_dense=n_dense, n_model=1024, feat_dim=feat_dim, n_proposal1=16)
    bg = b["bg"][:, None, :]
    f1, f2 = torch.cat([bg, b["dense_fm"]], 1), torch.cat([bg, b["dense_fo"]], 1)
    atten = oracle.feature_similarity(f1, f2, 0.1)
    oR, ot, os_, ost = oracle.fine_Rt(atten, b["dense_pm"], b["dense_po"], b["model"])
    R, t, s, st = pkg.model_utils.compute_fine_Rt_fused(f1.to(dev), f2.to(dev), b["dense_pm"].to(dev),
                                                        b["dense_po"].to(dev), b["model"].to(dev), 0.1, 0.15,
                                                        return_stages=True, precision=precision)
    for k in ("w1", "w2"):  # foreground decisions differ only at near ties
        assert (st[k].cpu() != ost[k]).float().mean().item() < 2e-3, k
    assert (st["row"].cpu() - ost["row"]).abs().max().item() < ROW_TOL[precision]
    assert (st["pred"].cpu() - ost["pred"]).abs().max().item() < PRED_TOL[precision]
    assert oracle.rotation_angle_deg(R.cpu(), oR).max().item() < 0.1   # north_star: <= 0.1 deg
    assert (t.cpu() - ot).abs().max().item() < 1e-4                    # <= 0.1 mm at 1 m radius
    assert (s.cpu() - os_).abs().max().item() < 2e-3


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("B,N1,N2,C,spt", [(2, 700, 300, 128, 0), (1, 1300, 1537, 64, 0), (3, 256, 257, 256, 1),
                                           (1, 1300, 600, 192, 4), (2, 1, 5, 64, 0),
                                           # C = 256: in fp16 mode the unrolled producer / issuer over the eight-slot ring
                                           # (slot set = parity of the running row step): ragged rows and columns, odd
                                           # step counts per task with several tasks per CTA pair (80 tasks on 74 pairs)
                                           (2, 1300, 1537, 256, 0), (10, 700, 2040, 256, 3), (5, 300, 2048, 256, 1)])
def test_fine_fused_ragged_shapes_and_task_splits(pkg, oracle, dev, B, N1, N2, C, spt, precision):
    """N1 != N2, sizes that are not multiples of the 256-wide tiles, every row-steps-per-task
    setting (0 = the host's choice): correlated random features so that the assignment is not flat."""
    import ctypes
    g = torch.Generator().manual_seed(N1 * 7 + N2)
    base = torch.randn(B, max(N1, N2) + 1, C, generator=g)
    f1 = (base[:, : N1 + 1] + 0.3 * torch.randn(B, N1 + 1, C, generator=g)).contiguous()
    f2 = (base[:, : N2 + 1] + 0.3 * torch.randn(B, N2 + 1, C, generator=g)).contiguous()
    p1 = (0.3 * torch.randn(B, N1, 3, generator=g)).contiguous()
    p2 = (0.3 * torch.randn(B, N2, 3, generator=g)).contiguous()
    model = (0.3 * torch.randn(B, 500, 3, generator=g)).contiguous()
    atten = oracle.feature_similarity(f1, f2, 0.1)
    oR, ot, os_, ost = oracle.fine_Rt(atten, p1, p2, model)
    R, t, s, st = pkg.model_utils.compute_fine_Rt_fused(f1.to(dev), f2.to(dev), p1.to(dev), p2.to(dev),
                                                        model.to(dev), 0.1, 0.15, return_stages=True,
                                                        precision=precision, steps_per_task=spt)
    for k in ("w1", "w2"):
        bad = (st[k].cpu() != ost[k]).float().sum().item()
        assert bad <= max(1.0, 2e-3 * ost[k].numel()), (k, bad)
    assert (st["row"].cpu() - ost["row"]).abs().max().item() < ROW_TOL[precision]
    if min(N1, N2) >= 256:  # a well-conditioned Procrustes problem
        assert (st["pred"].cpu() - ost["pred"]).abs().max().item() < PRED_TOL[precision]
        assert oracle.rotation_angle_deg(R.cpu(), oR).max().item() < 0.1
        assert (t.cpu() - ot).abs().max().item() < 1e-4
        # the score counts points within dis_thres: steps of 1 / N1, two borderline points on the small clouds
        assert (s.cpu() - os_).abs().max().item() < max(2e-3, 2.1 / N1)


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("n_unrelated", [0, 40, 700, 2048])
def test_fine_fused_two_pass_equals_three_pass(pkg, oracle, dev, n_unrelated, precision):
    """The default path merges the label pass with the weighted sums over ALL columns and takes the
    background columns out again with a short ragged pass; the original three full passes stay
    selectable.  Both must agree with each other and with the oracle for any number of background
    columns: none, fewer than one 256-wide strip, several strips, (nearly) all of them."""
    import ctypes
    g = torch.Generator().manual_seed(n_unrelated + 5)
    B, N, C = 2, 2048, 128
    base = torch.randn(B, N + 1, C, generator=g)
    f1 = (base + 0.3 * torch.randn(B, N + 1, C, generator=g)).contiguous()
    f2 = (base + 0.3 * torch.randn(B, N + 1, C, generator=g)).contiguous()
    if n_unrelated:  # observation points whose best match is the background token: they end up background
        idx = 1 + torch.randperm(N, generator=g)[:n_unrelated]
        f2[:, idx] = f1[:, 0:1] + 0.2 * torch.randn(B, n_unrelated, C, generator=g)
    p1 = (0.3 * torch.randn(B, N, 3, generator=g)).contiguous()
    p2 = (0.3 * torch.randn(B, N, 3, generator=g)).contiguous()
    model = (0.3 * torch.randn(B, 500, 3, generator=g)).contiguous()
    out = {}
    for passes in (3, 2):
        out[passes] = pkg.model_utils.compute_fine_Rt_fused(f1.to(dev), f2.to(dev), p1.to(dev), p2.to(dev),
                                                            model.to(dev), 0.1, 0.15, return_stages=True,
                                                            precision=precision, three_pass=passes == 3)
    (R3, t3, s3, st3), (R2, t2, s2, st2) = out[3], out[2]
    assert torch.equal(st3["w1"], st2["w1"]) and torch.equal(st3["w2"], st2["w2"])
    n_bg = int((st2["w2"] == 0).sum(1).max().item())
    assert n_bg >= min(n_unrelated, N) * 0.9, n_bg            # the case exercises what it claims
    assert (st3["row"] - st2["row"]).abs().max().item() < 2e-6
    live = st3["row"] > 1e-3                                   # rows that carry weight in the Procrustes
    assert ((st3["pred"] - st2["pred"]).abs().max(-1)[0] * live).max().item() < 1e-4
    if n_unrelated < N:
        assert oracle.rotation_angle_deg(R2.cpu(), R3.cpu()).max().item() < 1e-3
        atten = oracle.feature_similarity(f1, f2, 0.1)
        oR, ot, os_, ost = oracle.fine_Rt(atten, p1, p2, model)
        assert (st2["row"].cpu() - ost["row"]).abs().max().item() < ROW_TOL[precision]
        assert oracle.rotation_angle_deg(R2.cpu(), oR).max().item() < 0.1
        assert (t2.cpu() - ot).abs().max().item() < 1e-4
