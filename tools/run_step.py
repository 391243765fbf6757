"""A few graph replays of the stage and nothing else (the command ncu captures single kernels from)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device("cuda:0")
B = int(os.environ.get("B", 16))
stage = pkg.stage.GeometricStage(B, dev)
stage.load(pkg.synth.make_batch(list(range(B))))
for _ in range(int(os.environ.get("N", 3))):
    stage.run()
torch.cuda.synchronize()
print("ok", stage.records[0, :4].tolist())
