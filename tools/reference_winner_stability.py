#!/usr/bin/env python
"""Is the reference's "selected hypothesis index" a function of its inputs and seed?

Runs the reference's CPU arithmetic for compute_coarse_Rt (oracle/cpu_stage.py: the reference's own pointnet2
extension + the restatement of utils/model_utils.py that tests/test_oracle_cpu.py pins bit for bit on the live
reference) on the bench instances under different Intel MKL code paths (MKL_CBWR = AUTO / AVX2 / SSE4_2 ...: the
branch MKL takes on hosts with / without AVX-512) in separate processes and compares, per instance, the selected
proposal id, the coarse pose and the final (fine) pose.

TEST INFRASTRUCTURE (drives oracle/); writes profiles/r2_reference_winner_stability.json.

  python tools/reference_winner_stability.py [--instances 16] [--modes AUTO,AVX2,SSE4_2]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "oracle"))
import __graft_entry__ as e
pkg = e.load_package(); e.load_oracle()
import cpu_stage
torch.set_num_threads(min(8, os.cpu_count() or 1))
ids = list(range(%(n)d))
R, t, s, st = cpu_stage.run_stage_cpu(pkg.synth.make_batch(ids), faithful=False, chunk=4)
posed = torch.gather(st["coarse_posed"], 1, st["coarse_scores"].argmax(1, keepdim=True))[:, 0]
json.dump({"pid": st["proposal_id"].tolist(), "winner_well_posed": [bool(x) for x in posed],
           "R": R.tolist(), "t": t.tolist(), "score": s.tolist(), "init_R": st["init_R"].tolist(),
           "init_t": st["init_t"].tolist()}, sys.stdout)
'''


def run_mode(mode, n):
    env = dict(os.environ, MKL_CBWR=mode)
    out = subprocess.run([sys.executable, "-c", WORKER % {"root": ROOT, "n": n}], env=env, check=True,
                         capture_output=True, text=True).stdout
    return json.loads(out[out.index("{"):])


def compare(a, b):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pem_oracle as O
    n = len(a["pid"])
    same = [x == y for x, y in zip(a["pid"], b["pid"])]
    Ra, Rb = torch.tensor(a["R"]), torch.tensor(b["R"])
    ia, ib = torch.tensor(a["init_R"]), torch.tensor(b["init_R"])
    return {"selected_index_identical": "%d/%d" % (sum(same), n),
            "coarse_pose_max_dR_deg": O.rotation_angle_deg(ia, ib).max().item(),
            "final_pose_max_dR_deg": O.rotation_angle_deg(Ra, Rb).max().item(),
            "final_pose_max_dt": (torch.tensor(a["t"]) - torch.tensor(b["t"])).abs().max().item(),
            "final_score_max_diff": (torch.tensor(a["score"]) - torch.tensor(b["score"])).abs().max().item()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=16)
    ap.add_argument("--modes", default="AUTO,AVX2,SSE4_2")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_reference_winner_stability.json"))
    args = ap.parse_args()
    modes = args.modes.split(",")
    runs = {m: run_mode(m, args.instances) for m in modes}
    import torch
    rep = {"what": "reference CPU arithmetic of the whole stage on bench instances 0..%d, one process per MKL code "
                   "path (MKL_CBWR); torch %s" % (args.instances - 1, torch.__version__),
           "selected_proposal_id": {m: runs[m]["pid"] for m in modes},
           "winner_is_well_posed_triplet": {m: runs[m]["winner_well_posed"] for m in modes},
           "pairs": {"%s vs %s" % (modes[0], m): compare(runs[modes[0]], runs[m]) for m in modes[1:]}}
    with open(args.out, "w") as f:
        json.dump(rep, f, indent=1)
    print(json.dumps(rep["pairs"], indent=1))


if __name__ == "__main__":
    main()
