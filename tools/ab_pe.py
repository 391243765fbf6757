"""Times the PE grouping kernels at the stage's shape (16 clouds of 2048 points per call, both scales)."""
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, fn, k in s.launches():
    if not (name.startswith("pe_") ):
        continue
    ms = []
    with torch.cuda.stream(s.stream):
        for _ in range(3): pkg._lib.check(fn(st), name)
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s.stream); pkg._lib.check(fn(st), name); e1.record(s.stream)
            s.stream.synchronize()
            ms.append(e0.elapsed_time(e1))
    print("%-40s %.1f us" % (name, sorted(ms)[len(ms) // 2] * 1e3))
