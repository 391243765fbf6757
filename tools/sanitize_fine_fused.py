"""compute-sanitizer target: the fused fine stage at small ragged shapes, both arithmetics; C = 256 takes the unrolled
producer / issuer in fp16 mode (several tasks per CTA pair with odd step counts), C = 128 the generic loops.
  compute-sanitizer --tool memcheck python tools/sanitize_fine_fused.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(0)
for B, N1, N2, C, spt in ((2, 300, 513, 256, 1), (10, 700, 2040, 256, 3), (1, 130, 257, 128, 0)):
    base = torch.randn(B, max(N1, N2) + 1, C, generator=gen)
    f1 = (base[:, : N1 + 1] + 0.3 * torch.randn(B, N1 + 1, C, generator=gen)).contiguous().to(dev)
    f2 = (base[:, : N2 + 1] + 0.3 * torch.randn(B, N2 + 1, C, generator=gen)).contiguous().to(dev)
    p1 = (0.3 * torch.randn(B, N1, 3, generator=gen)).to(dev)
    p2 = (0.3 * torch.randn(B, N2, 3, generator=gen)).to(dev)
    model = (0.3 * torch.randn(B, 500, 3, generator=gen)).to(dev)
    out = {}
    for precision in ("fp16", "split_bf16"):
        R, t, s = pkg.model_utils.compute_fine_Rt_fused(f1, f2, p1, p2, model, 0.1, 0.15, precision=precision, steps_per_task=spt)
        out[precision] = (R.cpu(), t.cpu(), s.cpu())
    dR = (out["fp16"][0] - out["split_bf16"][0]).abs().max().item()
    assert dR < 5e-3, dR
torch.cuda.synchronize()
print("SANITIZE_TARGET_OK")
