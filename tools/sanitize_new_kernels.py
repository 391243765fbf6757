"""compute-sanitizer target: the kernels rewritten in round 2 at small / ragged shapes (memcheck is slow).
  compute-sanitizer --tool memcheck python tools/sanitize_new_kernels.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(0)
mu, pu, ext = pkg.model_utils, pkg.pointnet2_utils, pkg.pointnet2_ext
# scoring: fixed (Nm = 1024, 2048) and generic (Nm = 300, 1500) forms, several batches per CTA, ragged foreground
for Nm, H, n1, nfg in ((1024, 90, 196, 150), (2048, 20, 64, 64), (300, 70, 50, 7), (1500, 33, 256, 256)):
    q, _ = torch.linalg.qr(torch.randn(2, H, 3, 3, generator=gen))
    R, t = q.contiguous().to(dev), (0.1 * torch.randn(2, H, 3, generator=gen)).to(dev)
    pts, model = (0.4 * torch.randn(2, n1, 3, generator=gen)).to(dev), (0.4 * torch.randn(2, Nm, 3, generator=gen)).to(dev)
    w1 = torch.zeros(2, n1)
    w1[:, :nfg] = 1.0
    a = mu.score_hypotheses(R, t, pts, w1.to(dev), model, kernel="tensor")
    b = mu.score_hypotheses(R, t, pts, w1.to(dev), model, kernel="simt")
    assert torch.allclose(a, b, rtol=1e-4), (Nm, (a - b).abs().max())
# PE grouping through the masks: N = 2048 (fixed form of the filter), 1000 and 33 (generic), dense cloud (long lists)
for N, scale in ((2048, 1.0), (1000, 1.0), (33, 1.0), (512, 0.05)):
    xyz = (scale * torch.randn(2, N, 3, generator=gen).clamp(-2, 2) * 0.4).to(dev).contiguous()
    masks = pu.pe_masks(xyz)
    g1, g2 = pu.pe_group2_masked(xyz, masks)
    h1, h2 = pu.pe_group2(xyz)
    assert torch.equal(g1, h1) and torch.equal(g2, h2), N
# gather / group rewrites: aligned and ragged
for B, C, N, M, ns in ((2, 3, 2048, 196, 32), (1, 7, 100, 33, 5), (2, 256, 2049, 196, 64), (1, 5, 60000, 77, 3)):
    f = torch.randn(B, C, N, generator=gen).to(dev)
    idx = torch.randint(0, N, (B, M), generator=gen, dtype=torch.int32).to(dev)
    out = ext.gather_points(f, idx)
    assert torch.equal(out, torch.gather(f, 2, idx.long().unsqueeze(1).expand(-1, C, -1)))
    gi = torch.randint(0, N, (B, M, ns), generator=gen, dtype=torch.int32).to(dev)
    out = ext.group_points(f, gi)
    want = torch.gather(f, 2, gi.long().reshape(B, 1, -1).expand(-1, C, -1)).reshape(B, C, M, ns)
    assert torch.equal(out, want)
torch.cuda.synchronize()
print("SANITIZE_TARGET_OK")
