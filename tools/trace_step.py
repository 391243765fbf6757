# NOTE: uses pem_dev_* tuning knobs: build the library with `make -C openvino-sam-6d_b200/csrc clean all EXTRA=-DPEM_DEV_KNOBS`
# (the shipped library keeps these as compile-time constants).
"""Kernel timeline of graph-replayed steps (torch profiler / CUPTI): per-kernel warm durations and
the idle gaps between consecutive kernels on the critical path."""
import os, sys, json
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device("cuda:0")
B = int(os.environ.get("B", 16))
if os.environ.get("PDL") == "0":  # plain launches instead of programmatic dependent ones
    import ctypes
    ctypes.CDLL(pkg._lib.LIB_PATH).pem_dev_set_pdl(0)
stage = pkg.stage.GeometricStage(B, dev)
stage.load(pkg.synth.make_batch(list(range(B))))
for _ in range(5):
    stage.run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(4):
        stage.run()
    torch.cuda.synchronize()
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "trace_step.json")
os.makedirs(os.path.dirname(path), exist_ok=True)
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ev.sort(key=lambda e: e["ts"])
# last replay = last len/4 events
n = len(ev) // 4
last = ev[-n:]
t0 = last[0]["ts"]
busy_end = t0
gap_total = 0.0
for e in last:
    gap = e["ts"] - busy_end
    if gap > 0:
        gap_total += gap
    print("%8.1f  +%6.1f us  dur %7.1f  stream %s  %s" % (e["ts"] - t0, max(gap, 0), e["dur"], e["args"].get("stream"), e["name"][:60]))
    busy_end = max(busy_end, e["ts"] + e["dur"])
print("step span %.1f us, idle gaps %.1f us, kernels %d" % (busy_end - t0, gap_total, n))
