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
