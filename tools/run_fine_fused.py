"""Run the fused fine stage a few times at the default shapes (ncu target)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device("cuda:0")
B = int(os.environ.get("B", 16))
base = pkg.synth.make_batch(list(range(4)))
b = {k: v.to(dev) for k, v in pkg.synth.replicate_batch(base, B // 4).items()}
bg = b["bg"][:, None, :]
f1, f2 = torch.cat([bg, b["dense_fm"]], 1).contiguous(), torch.cat([bg, b["dense_fo"]], 1).contiguous()
for _ in range(int(os.environ.get("ITERS", 3))):
    R, t, s = pkg.model_utils.compute_fine_Rt_fused(f1, f2, b["dense_pm"], b["dense_po"], b["model"], 0.1, 0.15,
                                                    precision=os.environ.get("PRECISION", "fp16"))
torch.cuda.synchronize()
print("score", s[:4].tolist())
