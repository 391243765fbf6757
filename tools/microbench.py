"""Per-kernel microbenchmarks with roofline fractions: BASELINE.json configs[3] (hypothesis-
scoring sweep) and configs[4] (pointnet2 ops on 64 clouds x 16k points -> 2048 samples), plus the
default-shape kernels of the stage.  Writes gpurun_out/microbench_<tag>.json.
CUDA events, 3 warm-ups, L2 flushed between timed launches."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
O = g.load_oracle()
mu, pu, ext = pkg.model_utils, pkg.pointnet2_utils, pkg.pointnet2_ext
dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
HBM = peaks["hbm_gbs"]
FP32_OPS = 148 * 128 * 1.965e9 / 1e12  # T lane-ops/s (one op per lane per cycle; FMA counts once)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {"peaks": {"hbm_gbs": HBM, "fp32_lane_tops": FP32_OPS, "bf16_tflops": peaks["bf16_tflops"]}, "rows": []}


def timeit(fn, iters=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def row(name, ms, **kw):
    r = {"kernel": name, "ms": ms}
    r.update(kw)
    out["rows"].append(r)
    print(json.dumps(r), flush=True)


def ellipsoid_clouds(B, N, seed):
    gen = torch.Generator().manual_seed(seed)
    u = torch.randn(B, N, 3, generator=gen)
    u = u / u.norm(dim=2, keepdim=True)
    ax = 0.4 + 0.6 * torch.rand(B, 1, 3, generator=gen)
    p = u * ax
    return (p / p.norm(dim=2).amax(1).view(B, 1, 1)).contiguous()


# ------------------------------------------------------------------ configs[4]: pointnet2 microbench
B, N, M = 64, 16384, 2048
xyz = ellipsoid_clouds(B, N, 1).to(dev)
ms = timeit(lambda: ext.furthest_point_sampling(xyz, M))
flop = B * (M - 1) * N * 10
row("fps 64x16384->2048", ms, gflops=flop / ms / 1e6, frac_fp32_lane_peak=flop / ms / 1e9 / FP32_OPS,
    streaming_equiv_gbs=B * (M - 1) * N * 20 / ms / 1e6, frac_hbm_streaming_equiv=B * (M - 1) * N * 20 / ms / 1e6 / HBM)
idx = ext.furthest_point_sampling(xyz, M)
chk = O.fps(xyz[:2].cpu().numpy(), M)
row("fps bit-exact vs oracle (2 clouds)", 0.0, equal=bool((idx[:2].cpu().numpy() == chk).all()))
centres = pu.gather_rows(xyz, idx)
for r, ns in ((0.1, 32), (0.2, 64)):
    ms = timeit(lambda: ext.ball_query(centres, xyz, r, ns))
    flop = B * M * N * 9
    row("ball_query 64x(2048 centres,16384 pts) r=%.1f ns=%d" % (r, ns), ms, gflops=flop / ms / 1e6,
        frac_fp32_lane_peak=flop / ms / 1e9 / FP32_OPS)
bq = ext.ball_query(centres, xyz, 0.2, 64)
want = O.ball_query(centres[:1].cpu().numpy(), xyz[:1].cpu().numpy(), 0.2, 64)
row("ball_query bit-exact vs oracle (1 cloud)", 0.0, equal=bool((bq[:1].cpu().numpy() == want).all()))
for C in (3, 256):
    feat = torch.randn(B, C, N, device=dev)
    ms = timeit(lambda: ext.group_points(feat, bq))
    byts = B * (4 * M * 64 + 8 * C * M * 64)
    row("group_points C=%d idx (2048,64)" % C, ms, gbs=byts / ms / 1e6, frac_hbm=byts / ms / 1e6 / HBM)
    del feat
feat = torch.randn(B, 256, N, device=dev)
ms = timeit(lambda: ext.gather_points(feat, idx))
byts = B * (4 * M + 8 * 256 * M)
row("gather_points C=256 -> 2048", ms, gbs=byts / ms / 1e6, frac_hbm=byts / ms / 1e6 / HBM)
featr = feat.transpose(1, 2).contiguous()
ms = timeit(lambda: pu.gather_rows(featr, idx))
row("gather_rows (point-major) C=256 -> 2048", ms, gbs=byts / ms / 1e6, frac_hbm=byts / ms / 1e6 / HBM)
del feat, featr, xyz, bq

# ------------------------------------------------------------------ default-shape stage kernels, B = 16
B = 16
base = pkg.synth.make_batch(list(range(4)))
b = {k: v.to(dev) for k, v in pkg.synth.replicate_batch(base, B // 4).items()}
pm, po = b["dense_pm"], b["dense_po"]
both = torch.cat([pm, po], 0).contiguous()
ms = timeit(lambda: ext.furthest_point_sampling(both, 196))
flop = 2 * B * 195 * 2048 * 10
row("fps 32x2048->196", ms, gflops=flop / ms / 1e6, frac_fp32_lane_peak=flop / ms / 1e9 / FP32_OPS)
ms = timeit(lambda: pu.pe_group2(both))
flop = 2 * B * 2048 * 2048 * 9
byts = 2 * B * 6 * 2048 * 96 * 4
row("pe_group2 32 clouds (both PE scales)", ms, gflops=flop / ms / 1e6, frac_fp32_lane_peak=flop / ms / 1e9 / FP32_OPS,
    out_gbs=byts / ms / 1e6)

masks = pu.pe_masks(both)
ms_f = timeit(lambda: pu.pe_masks(both))
ms_g = timeit(lambda: pu.pe_group2_masked(both, masks))
row("pe_masks 32 clouds (tensor-core candidate filter)", ms_f, pairs_per_sec=2 * B * 2048 * 2048 / ms_f * 1e3)
row("pe_group2_masked 32 clouds (exact re-test + output)", ms_g, out_gbs=byts / ms_g / 1e6)

# ------------------------------------------------------------------ SURVEY 8f "next" rows
# rank 2: template-side FPS at object load, 1 x 210000 -> 2048 (feature_extraction.py:152-158)
big = ellipsoid_clouds(1, 210000, 5).to(dev)
ms = timeit(lambda: ext.furthest_point_sampling(big, 2048), iters=3, warm=1)
flop = 2047 * 210000 * 10
row("fps 1x210000->2048 (object load, 16-CTA cluster)", ms, gflops=flop / ms / 1e6,
    frac_fp32_lane_peak=flop / ms / 1e9 / FP32_OPS)
# rank 3: GeometricStructureEmbedding.get_embedding_indices on the sparse points (transformer.py:304-341)
sp32 = torch.randn(2 * B, 197, 3, device=dev) * 0.4
ms = timeit(lambda: pkg.transformer.get_embedding_indices(sp32))
row("geo_embedding_indices 32x197 (d (N,N) + a (N,N,3))", ms, out_gbs=2 * B * 197 * 197 * 16 / ms / 1e6)
import time as _time
cpu_pts = sp32[:4].cpu()
t0 = _time.perf_counter()
O.geo_embedding_indices(cpu_pts)
row("geo_embedding_indices reference arithmetic on CPU, 4x197 (torch, all host threads)",
    (_time.perf_counter() - t0) * 1e3)

# rank 4: instance preparation of one 480x640 RGB-D frame with 16 detections (get_test_data's loop)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_instance_prep import _scene  # noqa: E402
dm, mk, sc_ = _scene(1, 480, 640, 16)
Kc = np.array([[572.4114, 0, 325.2611], [0, 573.57043, 242.04899], [0, 0, 1]])
d_dm, d_mk = torch.from_numpy(dm).to(dev), torch.from_numpy(mk).to(dev)
def _prep():
    return pkg.data_utils.prepare_instances(d_dm, Kc, d_mk, sc_, 0.21, 2048, 224, rng=np.random.RandomState(0))
for _ in range(3):
    res = _prep()
torch.cuda.synchronize()
t0 = _time.perf_counter()
for _ in range(10):
    res = _prep()
torch.cuda.synchronize()
ms = (_time.perf_counter() - t0) * 100
_L = pkg._lib.lib()
_nb = _L.pem_instance_scratch_bytes(16, 480, 640)
_scr, _inf = torch.empty(_nb, dtype=torch.uint8, device=dev), torch.empty((16, 8), dtype=torch.int32, device=dev)
ms_k = timeit(lambda: _L.pem_instance_candidates(d_dm.data_ptr(), d_mk.data_ptr(), 16, 480, 640, 572.4114, 573.57043,
                                                 325.2611, 242.04899, 1, 0.252, _inf.data_ptr(), _scr.data_ptr(), _nb,
                                                 None))
row("instance_candidates_kernel<f64> 480x640, 16 detections (one CTA each; largest mask %d points)"
    % int(_inf[:, 1].max().item()), ms_k)
row("prepare_instances 480x640, 16 detections (%d kept; wall clock incl. the host draw and its sync)" % len(res["kept"]), ms)
t0 = _time.perf_counter()
O.prepare_instances(dm, Kc, mk, sc_, 0.21, 2048, 224, np.random.RandomState(0), "f64")
row("prepare_instances reference arithmetic on CPU (numpy, oracle's sequential mean in a Python loop)",
    (_time.perf_counter() - t0) * 1e3)
t0 = _time.perf_counter()
wp = O.point_cloud_from_depth(dm, Kc, "f64")
for i_ in range(16):  # the reference's own numpy lines per detection (timing only)
    m_ = np.logical_and(mk[i_] > 0, dm > 0)
    if m_.sum() <= 32:
        continue
    y1_, y2_, x1_, x2_ = O.mask_bbox(m_)
    ch_ = m_[y1_:y2_, x1_:x2_].astype(np.float32).flatten().nonzero()[0]
    cl_ = wp.copy()[y1_:y2_, x1_:x2_, :].reshape(-1, 3)[ch_, :]
    fl_ = np.linalg.norm(cl_ - np.mean(cl_, axis=0)[None, :], axis=1) < 0.252
    ch_, cl_ = ch_[fl_], cl_[fl_]
    ci_ = np.random.choice(np.arange(len(ch_)), 2048, replace=len(ch_) <= 2048)
    O.resize_rgb_choose(ch_[ci_], [y1_, y2_, x1_, x2_], 224)
row("prepare_instances: the reference's numpy lines on the CPU, same frame", (_time.perf_counter() - t0) * 1e3)

# rank 1: PositionalEncoding front end (fine_point_matching.py:102-144): groups -> SharedMLP -> amax, fused
PE = pkg.positional_encoding
gen = torch.Generator().manual_seed(0)
def _branch():
    layers, ci = [], 6
    for co in (32, 64, 128):
        layers.append((torch.randn(co, ci, generator=gen) * (2.0 / ci) ** 0.5, 0.5 + torch.rand(co, generator=gen),
                       0.3 * torch.randn(co, generator=gen), 0.2 * torch.randn(co, generator=gen),
                       0.5 + torch.rand(co, generator=gen), 1e-5))
        ci = co
    return layers
L1, L2 = _branch(), _branch()
br1, br2 = [tuple(t.to(dev) for t in PE.fold_branch(L)) for L in (L1, L2)]
w3pe, b3pe = (torch.randn(256, 256, generator=gen) * 0.05).to(dev), torch.randn(256, generator=gen).to(dev)
pooled = torch.empty(2 * B, 2048, 256, device=dev)
i1, i2 = PE.pe_neighbours(both, masks)
ms1 = timeit(lambda: PE.pe_branch(both, i1, br1, pooled, 0))
ms2 = timeit(lambda: PE.pe_branch(both, i2, br2, pooled, 128))
fl = lambda ns: 2.0 * 2 * B * 2048 * ns * (6 * 32 + 32 * 64 + 64 * 128)
row("pe_mlp branch ns=32, 32 clouds (group+SharedMLP+amax fused, split-bf16 tcgen05)", ms1, algorithmic_tflops=fl(32) / ms1 / 1e9,
    executed_mma_tflops=3 * fl(32) / ms1 / 1e9)
row("pe_mlp branch ns=64, 32 clouds", ms2, algorithmic_tflops=fl(64) / ms2 / 1e9, executed_mma_tflops=3 * fl(64) / ms2 / 1e9)
ms = timeit(lambda: PE.positional_encoding(both, br1, br2, w3pe, b3pe, masks=masks))
row("positional_encoding 32 clouds end to end (neighbour lists + 2 branches + Linear 256)", ms)
def _torch_pe():  # the reference's formulation in plain PyTorch fp32 on the same GPU (grouped tensor + conv2d + bn + relu + amax)
    g1, g2 = pu.pe_group2_masked(both, masks)
    outs = []
    for grp, layers in ((g1, L1), (g2, L2)):
        x = grp
        for (w, gw, gb, rm, rv, eps) in layers:
            x = torch.nn.functional.conv2d(x, w.to(dev).reshape(w.shape[0], w.shape[1], 1, 1))
            x = torch.relu(torch.nn.functional.batch_norm(x, rm.to(dev), rv.to(dev), gw.to(dev), gb.to(dev), False, 0.0, eps))
        outs.append(torch.amax(x, dim=3))
    return torch.nn.functional.conv1d(torch.cat(outs, 1), w3pe[:, :, None], b3pe)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
ms = timeit(_torch_pe, iters=3, warm=2)
row("positional_encoding 32 clouds, plain PyTorch fp32 (grouped tensor materialised, cuDNN convs)", ms)

# ------------------------------------------------------------------ configs[3]: scoring sweep
sp, sf, _ = mu.sample_pts_feats(pm, b["dense_fm"], 196, True)
tp, tf, _ = mu.sample_pts_feats(po, b["dense_fo"], 196, True)
bg = b["bg"][:, None, :]
ac = mu.compute_feature_similarity(torch.cat([bg, sf], 1), torch.cat([bg, tf], 1), "cosine", 0.1)
model2 = pkg.synth.make_batch([0], n_model=2048)["model"].to(dev).repeat(B, 1, 1).contiguous()
for P1 in (6000, 12000, 24000, 36000, 48000, 60000):
    rand = torch.rand(B, 3 * P1, device=dev)
    _, _, st = mu.compute_coarse_Rt(ac, sp, tp, model2, P1, 300, rand=rand, return_stages=True)
    R, t, w1 = st["R"], st["t"], st["w1"]
    nfg = float(w1.sum().item()) / B
    ms = timeit(lambda: mu.score_hypotheses(R, t, sp, w1, model2), iters=3)
    pairs = B * P1 * nfg * 2048
    row("score_hypotheses sweep P=%d x Nm=2048 x B=16" % P1, ms, hyps_per_sec=B * P1 / ms * 1e3,
        lane_tops=pairs * 4 / ms / 1e9, frac_fp32_lane_peak=pairs * 4 / ms / 1e9 / FP32_OPS,
        mean_foreground_points=nfg)
    ms2 = timeit(lambda: mu.compute_coarse_Rt(ac, sp, tp, model2, P1, P1, rand=rand), iters=3)
    row("compute_coarse_Rt P1=P2=%d (sample+solve+select+score all)" % P1, ms2, hyps_per_sec=B * P1 / ms2 * 1e3)

tag = os.environ.get("TAG", "r1")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "microbench_%s.json" % tag), "w"), indent=1)
