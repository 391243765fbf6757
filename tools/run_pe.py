import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device("cuda:0")
base = pkg.synth.make_batch(list(range(4)))
b = {k: v.to(dev) for k, v in pkg.synth.replicate_batch(base, 4).items()}
xyz = b["dense_po"].contiguous()
for _ in range(3):
    o1, o2 = pkg.pointnet2_utils.pe_group2(xyz)
torch.cuda.synchronize()
print(o1.shape, o2.shape)
