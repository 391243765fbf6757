"""Per-op CUDA-event timings at the default shapes (development aid, not the bench)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
mu, pu, ext = pkg.model_utils, pkg.pointnet2_utils, pkg.pointnet2_ext
dev = torch.device("cuda:0")
B = int(os.environ.get("B", 16))
base = pkg.synth.make_batch(list(range(4)))
b = {k: v.to(dev) for k, v in pkg.synth.replicate_batch(base, B // 4).items()}


def timeit(name, fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%-42s %9.1f us" % (name, e0.elapsed_time(e1) / iters * 1e3), flush=True)


pm, po, fm, fo = b["dense_pm"], b["dense_po"], b["dense_fm"], b["dense_fo"]
timeit("fps 2048->196 (B=%d)" % B, lambda: ext.furthest_point_sampling(pm, 196))
both = torch.cat([pm, po], 0)
timeit("fps 2048->196 (2B=%d)" % (2 * B), lambda: ext.furthest_point_sampling(both, 196))
idx = ext.furthest_point_sampling(pm, 196)
timeit("gather_rows 256ch", lambda: pu.gather_rows(fm, idx))
fmT = fm.transpose(1, 2).contiguous()
timeit("gather_points (B,256,2048)", lambda: ext.gather_points(fmT, idx))
timeit("sample_pts_feats", lambda: mu.sample_pts_feats(pm, fm, 196))
cen = (po + 1e-8).contiguous()
timeit("ball_query r=.1 ns=32", lambda: ext.ball_query(cen, po, 0.1, 32))
timeit("ball_query r=.2 ns=64", lambda: ext.ball_query(cen, po, 0.2, 64))
poT = po.transpose(1, 2).contiguous()
timeit("query_and_group r=.1 ns=32", lambda: pu.query_and_group(po, cen, poT, 0.1, 32))
timeit("query_and_group r=.2 ns=64", lambda: pu.query_and_group(po, cen, poT, 0.2, 64))
bq = ext.ball_query(cen, po, 0.2, 64)
timeit("group_points C=3 ns=64", lambda: ext.group_points(poT, bq))
sp, sf, _ = mu.sample_pts_feats(pm, fm, 196, True)
tp, tf, _ = mu.sample_pts_feats(po, fo, 196, True)
bg = b["bg"][:, None, :]
fc1, fc2 = torch.cat([bg, sf], 1), torch.cat([bg, tf], 1)
timeit("similarity coarse 197", lambda: mu.compute_feature_similarity(fc1, fc2, "cosine", 0.1))
ac = mu.compute_feature_similarity(fc1, fc2, "cosine", 0.1)
timeit("compute_coarse_Rt 6000/300", lambda: mu.compute_coarse_Rt(ac, sp, tp, b["model"], 6000, 300, rand=b["rand"]))
L, st = pkg._lib.lib(), ext._stream()
_, _, stg = mu.compute_coarse_Rt(ac, sp, tp, b["model"], 6000, 300, rand=b["rand"], return_stages=True)
w1 = stg["w1"]; w2 = torch.empty_like(w1); cdf = stg["cdf"]
timeit("  coarse_assign", lambda: L.pem_coarse_assign(ac.data_ptr(), B, 196, 196, w1.data_ptr(), w2.data_ptr(), cdf.data_ptr(), st))
Rh, th, rs = stg["R"], stg["t"], stg["resid"]
timeit("  coarse_hypotheses", lambda: L.pem_coarse_hypotheses(cdf.data_ptr(), b["rand"].data_ptr(), sp.data_ptr(), tp.data_ptr(), B, 196, 196, 6000, Rh.data_ptr(), th.data_ptr(), rs.data_ptr(), None, st))
sel = stg["sel"]
timeit("  select_smallest", lambda: L.pem_select_smallest(rs.data_ptr(), B, 6000, 300, sel.data_ptr(), st))
timeit("  score_hypotheses 300 x Nm=1024", lambda: mu.score_hypotheses(Rh, th, sp, w1, b["model"], sel=sel))
timeit("  score_hypotheses 6000 x Nm=1024", lambda: mu.score_hypotheses(Rh, th, sp, w1, b["model"]))
ff1, ff2 = torch.cat([bg, fm], 1), torch.cat([bg, fo], 1)
timeit("similarity fine 2049 (fp32 simt)", lambda: mu.compute_feature_similarity(ff1, ff2, "cosine", 0.1), iters=5)
af = mu.compute_feature_similarity(ff1, ff2, "cosine", 0.1)
timeit("compute_fine_Rt (materialised)", lambda: mu.compute_fine_Rt(af, pm, po, b["model"]), iters=5)
R, t, s = mu.compute_fine_Rt(af, pm, po, b["model"])
print("score", s[:4].tolist())
