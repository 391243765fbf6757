"""Run the PE grouping kernels at the default shapes (ncu target)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device("cuda:0")
base = pkg.synth.make_batch(list(range(4)))
b = {k: v.to(dev) for k, v in pkg.synth.replicate_batch(base, 4).items()}
xyz = torch.cat([b["dense_pm"] - b["gt_t"][:, None, :], b["dense_po"]], 0).contiguous()
cen = (xyz + 1e-8).contiguous()
featT = xyz.transpose(1, 2).contiguous()
for _ in range(3):
    o1 = pkg.pointnet2_utils.query_and_group(xyz, cen, featT, 0.1, 32)
    o2 = pkg.pointnet2_utils.query_and_group(xyz, cen, featT, 0.2, 64)
torch.cuda.synchronize()
print(o1.shape, o2.shape)
