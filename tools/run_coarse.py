import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device("cuda:0")
mu = pkg.model_utils
base = pkg.synth.make_batch(list(range(4)))
b = {k: v.to(dev) for k, v in pkg.synth.replicate_batch(base, 4).items()}
for _ in range(3):
    sp, sf, _ = mu.sample_pts_feats(b["dense_pm"], b["dense_fm"], 196, True)
    tp, tf, _ = mu.sample_pts_feats(b["dense_po"], b["dense_fo"], 196, True)
    bg = b["bg"][:, None, :]
    ac = mu.compute_feature_similarity(torch.cat([bg, sf], 1), torch.cat([bg, tf], 1), "cosine", 0.1)
    R, t = mu.compute_coarse_Rt(ac, sp, tp, b["model"], 6000, 300, rand=b["rand"])
torch.cuda.synchronize()
print(R[0])
