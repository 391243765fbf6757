"""PE grouping at the default shapes: one-shot kernel and the two-phase pair (ncu / timing target)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device("cuda:0")
B = int(os.environ.get("B", 16))
b = pkg.synth.make_batch(list(range(B)))
xyz = b["dense_po"].to(dev).contiguous()
pu = pkg.pointnet2_utils
def timeit(name, fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    print("%-28s %8.1f us" % (name, e0.elapsed_time(e1) / iters * 1e3), flush=True)
masks = pu.pe_masks(xyz)
timeit("pe_group2 one-shot", lambda: pu.pe_group2(xyz))
timeit("pe_masks (tensor core)", lambda: pu.pe_masks(xyz))
timeit("pe_group2_masked", lambda: pu.pe_group2_masked(xyz, masks))
