# NOTE: uses pem_dev_* tuning knobs: build the library with `make -C openvino-sam-6d_b200/csrc clean all EXTRA=-DPEM_DEV_KNOBS`
# (the shipped library keeps these as compile-time constants).
"""Fused fine kernel: parity vs the oracle + timing, for several row-steps-per-task settings (0 = automatic)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package(); O = g.load_oracle()
mu = pkg.model_utils
dev = torch.device("cuda:0")
L = ctypes.CDLL(pkg._lib.LIB_PATH)
b = pkg.synth.make_batch([32, 36], n_dense=2048, n_model=1024, feat_dim=256, n_proposal1=16)
bg = b["bg"][:, None, :]
f1, f2 = torch.cat([bg, b["dense_fm"]], 1), torch.cat([bg, b["dense_fo"]], 1)
atten = O.feature_similarity(f1, f2, 0.1)
oR, ot, os_, ost = O.fine_Rt(atten, b["dense_pm"], b["dense_po"], b["model"])
gd = {k: v.to(dev) for k, v in b.items()}
f1d, f2d = f1.to(dev), f2.to(dev)
big = {k: v.to(dev) for k, v in pkg.synth.replicate_batch(pkg.synth.make_batch(list(range(4))), 4).items()}
bgb = big["bg"][:, None, :]
F1, F2 = torch.cat([bgb, big["dense_fm"]], 1).contiguous(), torch.cat([bgb, big["dense_fo"]], 1).contiguous()
for variant in (0, 8, 4, 2, 1):
    L.pem_dev_set_fine_steps_per_task(variant)
    got = mu.compute_feature_similarity(f1d, f2d, "cosine", 0.1, True, backend="tc").cpu()
    print("steps/task", variant, "similarity max err", (got - atten).abs().max().item())
    R, t, s, st = mu.compute_fine_Rt_fused(f1d, f2d, gd["dense_pm"], gd["dense_po"], gd["model"], 0.1, 0.15, return_stages=True)
    print("   w1 mismatch", (st["w1"].cpu() != ost["w1"]).float().mean().item(), "w2", (st["w2"].cpu() != ost["w2"]).float().mean().item(),
          "row err", (st["row"].cpu() - ost["row"]).abs().max().item(), "pred err", (st["pred"].cpu() - ost["pred"]).abs().max().item())
    print("   dR deg", O.rotation_angle_deg(R.cpu(), oR).max().item(), "dt", (t.cpu() - ot).abs().max().item(), "score", s.tolist(), os_.tolist())
    for _ in range(3):
        mu.compute_fine_Rt_fused(F1, F2, big["dense_pm"], big["dense_po"], big["model"], 0.1, 0.15)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        mu.compute_fine_Rt_fused(F1, F2, big["dense_pm"], big["dense_po"], big["model"], 0.1, 0.15)
    e1.record(); torch.cuda.synchronize()
    print("   fused fine B=16: %.1f us" % (e0.elapsed_time(e1) / 10 * 1e3))
L.pem_dev_set_fine_steps_per_task(0)
