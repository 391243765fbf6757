"""GPU parity of the whole geometric stage (stage.py, CUDA-graph and eager) against the
host stage in the reference's arithmetic (oracle/cpu_stage.py)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cpu_result(pkg, oracle):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_stage
    batch = pkg.synth.make_batch([40, 41])
    R, t, s, st = cpu_stage.run_stage_cpu(batch, faithful=False)
    return batch, R, t, s, st


@pytest.mark.parametrize("use_graph,fine_precision", [(False, "fp16"), (True, "fp16"), (True, "split_bf16")])
def test_stage_matches_cpu_reference(pkg, oracle, dev, cpu_result, use_graph, fine_precision):
    batch, oR, ot, os_, ost = cpu_result
    B = 2
    stage = pkg.stage.GeometricStage(B, dev, use_graph=use_graph, fine_precision=fine_precision)
    stage.load(batch)
    for _ in range(2):  # second run replays the captured graph
        stage.run()
    stage.stream.synchronize()
    R, t, s, pid = stage.result()
    # bit-exact integer/byte work
    np.testing.assert_array_equal(stage.fps_idx[0].cpu().numpy(), ost["idx_m"].numpy())
    np.testing.assert_array_equal(stage.fps_idx[1].cpu().numpy(), ost["idx_o"].numpy())
    for k, (r, ns) in enumerate(pkg.stage.PE_SCALES):  # template-cloud PE groups (pose independent)
        np.testing.assert_array_equal(stage.pe_out[k][B:].cpu().numpy(), ost["pe"][2 + k].numpy())
    for blk in range(3):
        np.testing.assert_array_equal(stage.sparse_fine[blk][0].cpu().numpy(), ost["fine_sparse"][2 * blk].numpy())
        np.testing.assert_array_equal(stage.sparse_fine[blk][1].cpu().numpy(), ost["fine_sparse"][2 * blk + 1].numpy())
    # poses: coarse winner identical (or equal-score twin), final pose within north_star tolerance
    ang0 = oracle.rotation_angle_deg(stage.init_R.view(B, 3, 3).cpu(), ost["init_R"])
    assert ang0.max().item() < 0.1, ang0
    assert oracle.rotation_angle_deg(R.cpu(), oR).max().item() < 0.1      # <= 0.1 deg
    assert (t.cpu() - ot).abs().max().item() < 1e-4                       # <= 0.1 mm at 1 m radius
    assert (s.cpu() - os_).abs().max().item() < 2e-3
    rec = stage.records.cpu()
    assert torch.equal(rec[:, :9], R.reshape(B, 9).cpu()) and torch.equal(rec[:, 12], s.cpu())
    assert stage.kernels_per_step > 20


def test_pipelined_stage_with_half_precision_upload(pkg, oracle, dev):
    """The host-facing path of bench.py's `e2e` figure: PipelinedStage(upload="fp16") takes the caller's features in
    half precision (half the bytes on the link) and widens them on the device.  Parity is against the reference fed
    the SAME fp16-representable features (PipelinedStage.as_uploaded); two different batches alternate over the two
    replicas so that the per-batch background token and instance data really follow the upload."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import cpu_stage
    B = 2           # reference shapes (the north_star bars are stated for them)
    kw, gen = {}, {}
    batches = [pkg.synth.make_batch(list(range(g0, g0 + B)), **gen) for g0 in (100, 200)]
    pipe = pkg.stage.PipelinedStage(B, dev, depth=2, upload="fp16", **kw)
    pins = [pipe.pin_inputs(b) for b in batches]
    assert pins[0]["feats"].dtype == torch.float16
    with pytest.raises(RuntimeError):
        pkg.stage.PipelinedStage(B, dev, depth=1, upload="fp32", **kw).submit(pins[0])   # dtype must match the mode
    got = []
    for k in range(4):
        if len(pipe.inflight) == len(pipe.stages):      # (submit() would collect -- and drop -- the oldest itself)
            got.append(pipe.collect().clone())
        pipe.submit(pins[k % 2])
    while pipe.inflight:
        got.append(pipe.collect().clone())
    assert len(got) == 4 and torch.equal(got[0], got[2]) and torch.equal(got[1], got[3])
    for k in range(2):
        ref_in = pkg.stage.PipelinedStage.as_uploaded(batches[k], "fp16")
        oR, ot, os_, _ = cpu_stage.run_stage_cpu(ref_in, faithful=False)
        rec = got[k]
        R, t, s = rec[:, :9].reshape(B, 3, 3), rec[:, 9:12], rec[:, 12]
        assert oracle.rotation_angle_deg(R, oR).max().item() < 0.1
        assert (t - ot).abs().max().item() < 1e-4
        assert (s - os_).abs().max().item() < 2e-3
