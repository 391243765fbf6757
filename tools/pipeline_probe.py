"""Throughput of the stage with DEPTH batches in flight (one GeometricStage replica, streams and CUDA graph per
batch in flight, device-resident inputs): does the latency-bound coarse chain of one batch hide under the bulk
kernels of another?  python tools/pipeline_probe.py [depth ...]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device("cuda:0")
B = 16
depths = [int(a) for a in sys.argv[1:]] or [1, 2, 3]
for depth in depths:
    stages = []
    for d in range(depth):
        s = pkg.stage.GeometricStage(B, dev)
        s.load(pkg.synth.make_batch(list(range(d * B, (d + 1) * B))))
        stages.append(s)
    for _ in range(3):
        for s in stages:
            s.run()
    torch.cuda.synchronize()
    K = 300
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream(dev)
    for s in stages:
        s.stream.wait_stream(cur)
    e0.record(cur)
    for s in stages:
        s.stream.wait_event(e0)
    for k in range(K):
        stages[k % depth].run()
    for s in stages:
        cur.wait_stream(s.stream)
    e1.record(cur)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("depth %d: %.1f us per step, %.0f instances/s" % (depth, 1e3 * ms / K, K * B / (ms / 1e3)))
    del stages
    torch.cuda.empty_cache()
