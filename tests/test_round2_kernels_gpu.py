"""Edge shapes of the kernels rewritten in round 2 (score_tc fixed / generic forms, the tensor-core candidate filter
and the warp-per-centre re-test incl. lists longer than one round, gather / group), against the kernels they
replaced or plain torch indexing.  (compute-sanitizer is closed on the GPU pool; tools/sanitize_new_kernels.py is the
same body as a stand-alone target.)"""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("Nm,H,n1,nfg", [(1024, 90, 196, 150), (2048, 20, 64, 64), (300, 70, 50, 7),
                                           (1500, 33, 256, 256), (896, 300, 196, 1), (897, 41, 196, 152)])
def test_scoring_forms_agree(pkg, dev, Nm, H, n1, nfg):
    gen = torch.Generator().manual_seed(Nm + H)
    mu = pkg.model_utils
    q, _ = torch.linalg.qr(torch.randn(2, H, 3, 3, generator=gen))
    R, t = q.contiguous().to(dev), (0.1 * torch.randn(2, H, 3, generator=gen)).to(dev)
    pts, model = (0.4 * torch.randn(2, n1, 3, generator=gen)).to(dev), (0.4 * torch.randn(2, Nm, 3, generator=gen)).to(dev)
    w1 = torch.zeros(2, n1)
    w1[:, torch.randperm(n1, generator=gen)[:nfg]] = 1.0
    a = mu.score_hypotheses(R, t, pts, w1.to(dev), model, kernel="tensor")
    b = mu.score_hypotheses(R, t, pts, w1.to(dev), model, kernel="simt")
    rel = ((a - b).abs() / b.abs().clamp_min(1e-6)).max().item()
    print("Nm %d nfg %d: max relative difference tensor vs simt %.2e" % (Nm, nfg, rel))
    # (with a handful of points the score is 1 / a few small distances and carries the expanded form's cancellation
    # error undamped: d^2 to ~1e-7 absolute in either kernel)
    assert rel < (2e-5 if nfg >= 32 else 2e-3)
    assert torch.equal(a, mu.score_hypotheses(R, t, pts, w1.to(dev), model))       # auto = tensor for these shapes


@pytest.mark.parametrize("N,scale", [(2048, 1.0), (1000, 1.0), (33, 1.0), (512, 0.05), (4096, 0.6), (3000, 0.02)])
def test_masked_pe_grouping_equals_the_exact_scan(pkg, dev, N, scale):
    """scale 0.05 / 0.02: every point within both radii of every centre -- candidate lists of N entries (several
    rounds of the 256-entry list), rows full after the first few."""
    gen = torch.Generator().manual_seed(N)
    pu = pkg.pointnet2_utils
    xyz = (scale * torch.randn(2, N, 3, generator=gen).clamp(-2, 2) * 0.4).to(dev).contiguous()
    masks = pu.pe_masks(xyz)
    g1, g2 = pu.pe_group2_masked(xyz, masks)
    h1, h2 = pu.pe_group2(xyz)
    assert torch.equal(g1, h1) and torch.equal(g2, h2)


@pytest.mark.parametrize("B,C,N,M,ns", [(2, 3, 2048, 196, 32), (1, 7, 100, 33, 5), (2, 256, 2049, 196, 64),
                                         (1, 5, 60000, 77, 3), (2, 4, 16384, 2048, 16), (1, 1, 9, 1, 1)])
def test_gather_and_group_against_torch_indexing(pkg, dev, B, C, N, M, ns):
    gen = torch.Generator().manual_seed(N + M)
    ext = pkg.pointnet2_ext
    f = torch.randn(B, C, N, generator=gen).to(dev)
    idx = torch.randint(0, N, (B, M), generator=gen, dtype=torch.int32).to(dev)
    assert torch.equal(ext.gather_points(f, idx), torch.gather(f, 2, idx.long().unsqueeze(1).expand(-1, C, -1)))
    gi = torch.randint(0, N, (B, M, ns), generator=gen, dtype=torch.int32).to(dev)
    want = torch.gather(f, 2, gi.long().reshape(B, 1, -1).expand(-1, C, -1)).reshape(B, C, M, ns)
    assert torch.equal(ext.group_points(f, gi), want)
