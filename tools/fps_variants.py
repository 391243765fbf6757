# NOTE: uses pem_dev_* tuning knobs: build the library with `make -C openvino-sam-6d_b200/csrc clean all EXTRA=-DPEM_DEV_KNOBS`
# (the shipped library keeps these as compile-time constants).
import os, sys, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device("cuda:0")
ext = pkg.pointnet2_ext
L = ctypes.CDLL(pkg._lib.LIB_PATH)
base = pkg.synth.make_batch(list(range(4)))
b = pkg.synth.replicate_batch(base, 4)
both = torch.cat([b["dense_pm"], b["dense_po"]], 0).contiguous().to(dev)
ref = None
for v in (0, 1, 2, 3):
    L.pem_dev_set_fps_variant(v)
    for _ in range(3):
        idx = ext.furthest_point_sampling(both, 196)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        idx = ext.furthest_point_sampling(both, 196)
    e1.record(); torch.cuda.synchronize()
    if ref is None: ref = idx.clone()
    print("variant", v, "%.1f us" % (e0.elapsed_time(e1) / 20 * 1e3), "same", torch.equal(idx, ref))
L.pem_dev_set_fps_variant(0)
