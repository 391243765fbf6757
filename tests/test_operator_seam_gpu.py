"""The reference's operator seam (utils/model_utils.py:457-526: CustomSVD / CustomSVDu / CustomSVDv / CustomDet /
CustomSearchSorted) as standalone batched kernels, at the shapes of the reference's own op scripts
(ov_op_test/ov_test_custom_svd.py:31-33 and ov_test_custom_det.py: 42000 x 3 x 3, seed as in those scripts).

Singular vectors are unique only up to sign, so parity with LAPACK (torch.svd on the CPU = the oracle) is asserted on
what is unique: S, the reconstruction U diag(S) V^T, orthonormality, and the rotation the pose solvers build from the
factors, R = V diag(1,1,det(V U^T)) U^T (MU:415-420) -- tolerances stated next to each assert."""
import pytest
import torch

pytestmark = pytest.mark.gpu

N_MAT = 42000  # ov_test_custom_svd.py:31


def _R_from(U, V):
    D = torch.eye(3, dtype=U.dtype).repeat(U.shape[0], 1, 1)
    D[:, 2, 2] = torch.sign(torch.det(V @ U.transpose(1, 2)))
    return V @ D @ U.transpose(1, 2), D[:, 2, 2]


@pytest.fixture(scope="module")
def mats():
    torch.manual_seed(324)
    H = torch.randn(N_MAT, 3, 3)
    # rank-deficient rows as the coarse stage produces them: rank 2 (three centred points), rank 1
    # (a correspondence drawn twice) and H = 0 (one correspondence three times)
    a, b = torch.randn(200, 3, 1), torch.randn(200, 1, 3)
    H[:200] = a @ b
    c, d = torch.randn(200, 3, 2), torch.randn(200, 2, 3)
    H[200:400] = c @ d
    H[400:410] = 0.0
    return H


def test_custom_svd_42000(pkg, dev, mats):
    mu = pkg.model_utils
    U, S, V = (x.cpu() for x in mu.CustomSVD.apply(mats.to(dev)))
    eye = torch.eye(3)
    # orthonormal factors for EVERY input, rank-deficient or not (fp32 output rounding only)
    assert (U @ U.transpose(1, 2) - eye).abs().max().item() < 1e-6
    assert (V @ V.transpose(1, 2) - eye).abs().max().item() < 1e-6
    assert (S[:, :-1] >= S[:, 1:]).all() and (S >= 0).all()
    # reconstruction: entries of H are O(1), |H| up to ~8: 1e-5 absolute
    rec = U @ torch.diag_embed(S) @ V.transpose(1, 2)
    assert (rec - mats).abs().max().item() < 1e-5
    # singular values against LAPACK fp32 and against the fp64 decomposition
    oU, oS, oV = torch.svd(mats)
    assert (S - oS).abs().max().item() < 1e-5
    assert (S.double() - torch.linalg.svdvals(mats.double())).abs().max().item() < 2e-6
    # H = 0 -> U = V = I like LAPACK
    assert torch.equal(U[400:410], eye.expand(10, 3, 3)) and torch.equal(V[400:410], eye.expand(10, 3, 3))
    # the rotation of the pose solvers.  Its conditioning is sigma1 / (sigma2 + d*sigma3), d = the det sign:
    # LAPACK's own fp32 answer is that far from the fp64 one (measured on these matrices: err * gap < 8e-7), so
    # vs LAPACK fp32: 1e-5 where gap > 0.05, 3e-6/gap in general; vs the fp64 decomposition: fp32 rounding only.
    R, _ = _R_from(U, V)
    oR, _ = _R_from(oU, oV)
    xU, xS, xV = torch.svd(mats.double())
    xR, d = _R_from(xU, xV)
    gap = (xS[:, 1] + d * xS[:, 2]) / xS[:, 0].clamp_min(1e-300)
    err32 = (R - oR).abs().flatten(1).max(1)[0].double()
    err64 = (R.double() - xR).abs().flatten(1).max(1)[0]
    ok = gap > 1e-4
    ok[:410] = False
    assert ok.float().mean() > 0.95
    assert (err32[ok] * gap[ok]).max().item() < 3e-6
    assert err32[ok & (gap > 0.05)].max().item() < 1e-5
    assert (err64[ok] * gap[ok].clamp(max=1.0)).max().item() < 1e-6
    # rank 2 (rows 200..399): the rotation is still unique (the det correction picks the third axis)
    r2 = torch.arange(200, 400)
    g2 = xS[r2, 1] / xS[r2, 0]
    assert (err64[r2] * g2.clamp(max=1.0)).max().item() < 1e-5
    # every R is a proper rotation, also for rank 1 / 0
    assert (R @ R.transpose(1, 2) - eye).abs().max().item() < 2e-6
    assert (torch.det(R.double()) - 1).abs().max().item() < 1e-5
    # the single-output shells
    assert torch.equal(mu.CustomSVDu.apply(mats.to(dev)).cpu(), U)
    assert torch.equal(mu.CustomSVDv.apply(mats.to(dev)).cpu(), V)


def test_custom_det_42000(pkg, dev, mats):
    det = pkg.model_utils.CustomDet.apply(mats.to(dev)).cpu()
    want64 = torch.det(mats.double())
    scale = mats.abs().flatten(1).max(1)[0].double() ** 3
    assert ((det.double() - want64).abs() / scale.clamp_min(1e-30)).max().item() < 1e-6   # one fp32 rounding
    assert (det - torch.det(mats)).abs().max().item() < 1e-4                              # LAPACK fp32 LU, atol of ov_test_custom_det.py
    assert (det[400:410] == 0).all()
    # batched leading dimensions
    H4 = mats[:24].reshape(2, 3, 4, 3, 3).to(dev)
    assert torch.equal(pkg.model_utils.CustomDet.apply(H4).cpu().reshape(-1), det[:24])


@pytest.mark.parametrize("B,N,S", [(1, 38416, 18000), (4, 100, 7), (2, 1, 5), (3, 70000, 1000)])
def test_custom_searchsorted(pkg, dev, B, N, S):
    g = torch.Generator().manual_seed(N + S)
    inc = torch.rand(B, N, generator=g)
    inc[:, torch.rand(N, generator=g) < 0.3] = 0.0
    seq = torch.cumsum(inc.double(), 1)
    seq = (seq / seq[:, -1:].clamp_min(1e-30) * 0.97).float()
    val = torch.rand(B, S, generator=g)
    val[:, 0] = 1.0    # above the last entry -> N
    val[:, -1] = 0.0   # <= the first entry -> 0
    k = min(S, N) // 2
    if k:
        val[:, 1:1 + k] = seq[:, torch.randint(0, N, (k,), generator=g)]  # ties: the FIRST equal entry
    got = pkg.model_utils.CustomSearchSorted.apply(seq.to(dev), val.to(dev)).cpu()
    assert got.dtype == torch.int64
    assert torch.equal(got, torch.searchsorted(seq, val))
    assert (got[:, 0] == N).all() and (got[:, -1] == 0).all()
    one = pkg.model_utils.searchsorted(seq[0].to(dev), val[0].to(dev)).cpu()
    assert torch.equal(one, got[0])


def test_seam_shells_keep_the_reference_interface(pkg):
    """Same class names and ONNX symbolic op names as utils/model_utils.py:457-526."""
    mu = pkg.model_utils
    for name in ("CustomSearchSorted", "CustomSVD", "CustomSVDu", "CustomSVDv", "CustomDet"):
        cls = getattr(mu, name)
        assert issubclass(cls, torch.autograd.Function) and hasattr(cls, "symbolic")
