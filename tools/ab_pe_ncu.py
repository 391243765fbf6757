"""ncu target: the masked PE grouping of the stage's template cloud, a few launches."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device("cuda:0")
B = 16
s = pkg.stage.GeometricStage(B, dev)
s.load(pkg.synth.make_batch(list(range(B))))
s.run(); torch.cuda.synchronize()
st = s.stream.cuda_stream
for name, fn, k in s.launches():
    if name.startswith("pe_group2_masked(template"):
        with torch.cuda.stream(s.stream):
            for _ in range(3): pkg._lib.check(fn(st), name)
torch.cuda.synchronize()
print("ok")
