"""north_star asks for an "identical selected-hypothesis index".  tools/reference_winner_stability.py shows the
reference does not agree with ITSELF on that index: on the same inputs and seed, Intel MKL's AVX2 / SSE4.2 code
paths (what the reference runs on a host without AVX-512) pick a different proposal than the AVX-512 path on most
instances (profiles/r2_reference_winner_stability.json: 5/16 and 2/16 identical, coarse poses up to 17 degrees
apart), because ~23 % of the 6000 triplets are rank-1 and LAPACK's rotation for them is rounding noise
("degenerate triplets").  What IS a function of the inputs -- and what this repository asserts against
the reference on every bench instance (tests/test_bench_batch_gpu.py) -- is the final pose.  This CPU test keeps
the evidence live on a small sample: across MKL code paths the reference's final pose must agree within the
north_star bars whatever happens to the index."""
import os
import sys

import torch

from conftest import ROOT


def test_reference_final_pose_is_stable_across_mkl_code_paths_even_when_its_winner_is_not():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import reference_winner_stability as rws
    a, b = rws.run_mode("AUTO", 3), rws.run_mode("SSE4_2", 3)
    rep = rws.compare(a, b)
    print("reference vs itself (MKL AUTO vs SSE4_2, 3 instances):", rep)
    assert rep["final_pose_max_dR_deg"] < 0.1 and rep["final_pose_max_dt"] < 1e-4
    assert torch.isfinite(torch.tensor(a["R"])).all()
