"""CPU: host-side logic that needs no GPU -- synthetic instances, the pose-record sharding and
gather of the multi-GPU path (world_size 2 over gloo), bench.py's reference arm plumbing."""
import json
import os
import subprocess
import sys

import torch

from conftest import ROOT


def test_synth_is_deterministic_and_shaped(pkg):
    a = pkg.synth.make_batch([3, 4], n_dense=256, n_model=64, feat_dim=32, n_proposal1=10)
    b = pkg.synth.make_batch([3, 4], n_dense=256, n_model=64, feat_dim=32, n_proposal1=10)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert a["dense_pm"].shape == (2, 256, 3) and a["dense_fm"].shape == (2, 256, 32)
    assert a["model"].shape == (2, 64, 3) and a["rand"].shape == (2, 30) and a["bg"].shape == (2, 32)
    assert abs(a["dense_po"].norm(dim=2).max().item() - 1.0) < 1e-5          # radius-normalised template
    assert (torch.det(a["gt_R"]) - 1).abs().max() < 1e-5
    c = pkg.synth.make_batch([4], n_dense=256, n_model=64, feat_dim=32, n_proposal1=10)
    assert torch.equal(c["dense_pm"][0], a["dense_pm"][1])                   # one seed per instance id


def test_shard_plan(pkg):
    sh = pkg.parallel.shard_instances(10, 4)
    assert sh == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert pkg.parallel.shard_instances(256, 8) == [(32 * r, 32 * r + 32) for r in range(8)]


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
import __graft_entry__ as g
pkg = g.load_package()
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lo, hi = pkg.parallel.shard_instances(10, world)[rank]
rec = pkg.parallel.pack_records(torch.eye(3).repeat(hi - lo, 1, 1) * (rank + 1), torch.full((hi - lo, 3), float(rank)),
                                torch.arange(lo, hi).float() / 10, torch.arange(lo, hi, dtype=torch.int32) + 100,
                                torch.arange(lo, hi))
allrec = pkg.parallel.gather_records(rec, 10, world)
R, t, s, pid, iid = pkg.parallel.unpack_records(allrec)
assert iid.tolist() == list(range(10)), iid
assert pid.tolist() == [100 + i for i in range(10)]
assert torch.allclose(s, torch.arange(10).float() / 10)
assert t[:5].eq(0).all() and t[5:].eq(1).all()
assert R[7, 0, 0] == 2.0
# the same into a caller-owned table (what bench.py does every step)
table = torch.full((10, 16), -1.0)
assert pkg.parallel.gather_records(rec, 10, world, out=table) is table and torch.equal(table, allrec)
# uneven slices (4 + 3): padded collective, trimmed result
lo, hi = pkg.parallel.shard_instances(7, world)[rank]
rec7 = pkg.parallel.pack_records(torch.eye(3).repeat(hi - lo, 1, 1), torch.zeros(hi - lo, 3),
                                 torch.arange(lo, hi).float(), torch.arange(lo, hi, dtype=torch.int32),
                                 torch.arange(lo, hi))
all7 = pkg.parallel.gather_records(rec7, 7, world)
assert all7.shape == (7, 16) and pkg.parallel.unpack_records(all7)[4].tolist() == list(range(7))
if rank == 0:
    print("GATHER_OK")
dist.destroy_process_group()
"""


def test_pose_record_gather_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GATHER_OK" in out.stdout


def test_bench_reference_arm_contract():
    """--impl reference prints one JSON line with the contract's keys (tiny run)."""
    env = dict(os.environ, PEM_BENCH_CPU_INSTANCES="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "pem_instances_per_sec"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["value"] > 0 and line["cpu_baseline"]["cores"] >= 1


def test_dropin_registers_pointnet2_ext(pkg):
    """The reference's own pointnet2_utils.py imports cleanly on top of our `_ext` mirror
    (import-level check; the ops themselves need a GPU)."""
    import importlib
    import pytest
    ref = "/root/reference/SAM-6D/Pose_Estimation_Model/model/pointnet2"
    if not os.path.isdir(ref):
        pytest.skip("live reference not present")
    saved = {k: sys.modules.get(k) for k in ("pointnet2", "pointnet2._ext", "pointnet2_utils")}
    try:
        for k in saved:
            sys.modules.pop(k, None)
        ext = pkg.install_pointnet2_ext()
        assert sys.modules["pointnet2._ext"] is ext
        for name in ("gather_points", "gather_points_grad", "furthest_point_sampling", "three_nn",
                     "three_interpolate", "three_interpolate_grad", "ball_query", "group_points",
                     "group_points_grad"):                      # bindings.cpp:11-24
            assert callable(getattr(ext, name)), name
        sys.path.insert(0, ref)
        ref_utils = importlib.import_module("pointnet2_utils")
        assert ref_utils._ext is ext
        assert ref_utils.furthest_point_sample is not None and ref_utils.QueryAndGroup is not None
    finally:
        sys.path.remove(ref) if ref in sys.path else None
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v


def test_depth_dtype_is_checked_before_the_reinterpreting_view(pkg):
    """A non-16-bit depth image must raise instead of being read as a 16-bit image of another width."""
    import numpy as np
    import pytest
    for bad in (np.zeros((4, 6), np.int32), np.zeros((4, 6), np.float32), np.zeros((4, 6), np.uint8),
                np.zeros((4, 6), np.float16)):
        with pytest.raises(TypeError):
            pkg.data_utils.depth_to_metres(bad, 1.0)
    with pytest.raises(TypeError):
        pkg.data_utils.depth_to_metres(torch.zeros(4, 6, dtype=torch.int32), 1.0)
    with pytest.raises(ValueError):
        pkg.data_utils.depth_to_metres(np.zeros((2, 4, 6), np.uint16), 1.0)


def test_gather_rows_validates_its_tensors(pkg):
    """float16 / float64 / CPU feats must raise instead of being reinterpreted through a raw pointer."""
    import pytest
    idx = torch.zeros(1, 2, dtype=torch.int32)
    for feats in (torch.zeros(1, 4, 8, dtype=torch.float16), torch.zeros(1, 4, 8, dtype=torch.float64),
                  torch.zeros(1, 4, 8)):
        with pytest.raises(RuntimeError):
            pkg.pointnet2_utils.gather_rows(feats, idx)
        with pytest.raises(RuntimeError):
            pkg.model_utils.sample_pts_feats(torch.zeros(1, 4, 3), feats, 2)
