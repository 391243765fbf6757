#!/usr/bin/env python
"""bench.py -- PEM geometric-stage throughput on B200 (contract: see the task brief).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B]

A *step* is one pass of the geometric stage (openvino-sam-6d_b200/stage.py) over one batch
of B synthetic instances per GPU (default B = 16, BASELINE.json configs[1]: coarse 196 / fine
2048 points, 6000 -> 300 hypotheses, 1024 model points).  Prints ONE JSON line on rank 0.

  value        instances/s, whole job, inputs resident in HBM, CUDA-event time per step
               (L2 flushed by a 256 MB memset between steps, outside the timed events),
               max over ranks
  e2e          the same through the host-facing call: pinned host inputs -> H2D -> stage ->
               D2H of the 64-byte pose records, all inside the timed region
  roofline     the dominant kernel timed alone with CUDA events on its launch stream
  cpu_baseline the reference's CPU arithmetic (oracle/_ref pointnet2 extension + the torch-CPU
               restatement of model_utils.py) on a bounded sample, host cores stated
  --impl reference   only that CPU arm, same metric/config keys
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

METRIC = "pem_instances_per_sec"
UNIT = "instances/s"


def workload_config(args, world):
    return {
        "workload": "synthetic PEM geometric stage, configs[1]: %d instances/GPU, coarse 196 / fine "
                    "2048 points, C=256, 6000->300 hypotheses, 1024 model points" % args.batch,
        "instances_per_gpu": args.batch, "n_gpus": world, "n_dense": 2048, "n_sparse": 196,
        "feat_dim": 256, "n_model": 1024, "n_proposal1": 6000, "n_proposal2": 300,
        "parallelism": "instance-sharded dp%d, pose-record all_gather" % world,
        "l2": "flushed between timed steps (256 MB memset outside the event pair)",
        "fine_gemm": getattr(args, "fine_precision", "fp16") + " tcgen05 (fp32 accumulation; pose parity bars "
                     "0.1 deg / 1e-4 asserted for this mode in tests/test_tensorcore_gpu.py)",
    }


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 8 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 8 and r[2].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 8 and r[4 + k] == "Active" for r in self.rows)]
        # the busiest half of the samples = "under load"
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------- CPU arm
def cpu_arm(args, n_instances, reps):
    """Time the reference CPU arithmetic on `n_instances` instances, `reps` times."""
    odir = os.path.join(ROOT, "oracle")
    if odir not in sys.path:
        sys.path.insert(0, odir)
    import cpu_stage
    pkg = entry.load_package()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = pkg.synth.make_batch(list(range(n_instances)))
    ops = cpu_stage.pointnet2_ops()
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_stage.run_stage_cpu(batch, faithful=True, ops=ops)
        times.append(time.perf_counter() - t0)
    return {"seconds": times, "instances": n_instances, "cores": cores, "kind": ops.kind}


def run_reference(args, rank, world):
    if rank != 0:
        return
    n = int(os.environ.get("PEM_BENCH_CPU_INSTANCES", 2))
    r = cpu_arm(args, n, args.warmup + args.steps)
    timed = r["seconds"][args.warmup:]
    ms = 1e3 * sum(timed) / len(timed)
    value = n / (ms / 1e3)
    sample = "%d instances per step through the whole stage (reference sampler materialises " \
             "3.5 GB/instance, so the CPU arm is chunked to 2), %d timed steps" % (n, len(timed))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- GPU arm
def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist
    pkg = entry.load_package()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if os.environ.get("PEM_PDL") == "0":  # development: plain launches instead of programmatic dependent ones
        import ctypes
        ctypes.CDLL(pkg._lib.LIB_PATH).pem_dev_set_pdl(0)
    B, K, W = args.batch, args.steps, args.warmup
    ids = [rank * B + i for i in range(B)]
    host = pkg.synth.make_batch(ids)
    stage = pkg.stage.GeometricStage(B, dev, fine_precision=args.fine_precision)
    stage.instance_ids.copy_(torch.tensor(ids, dtype=torch.float32))
    stage.load(host)
    gathered = torch.empty((world * B, 16), dtype=torch.float32, device=dev)  # every rank's records

    def step_collective():
        # the only multi-GPU traffic: one all-gather of the 64-byte pose records (NCCL over NVLink),
        # straight from the stage's record buffer into a preallocated table
        if world > 1:
            with torch.cuda.stream(stage.stream):
                pkg.parallel.gather_records(stage.records, world * B, world, out=gathered)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- value: inputs resident, per-step CUDA events on the launch stream -------------
    clocks = ClockSampler(local_rank)
    clocks.start()
    for _ in range(max(W, 3)):
        stage.run()
        step_collective()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    wall0 = time.perf_counter()
    for k in range(K):
        with torch.cuda.stream(stage.stream):
            flush.zero_()
            ev[k][0].record(stage.stream)
        stage.run()
        step_collective()
        ev[k][1].record(stage.stream)
    barrier()
    wall = time.perf_counter() - wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)

    # ---- e2e: pinned host -> H2D -> stage -> D2H records, all inside the timed region --------
    # through the host-facing PipelinedStage (two replicas: copies of batch k+1 overlap batch k)
    pin = pkg.stage.PipelinedStage.pin_inputs(host)
    h2d = sum(v.numel() * v.element_size() for v in pin.values())
    d2h = B * 16 * 4
    pipe = pkg.stage.PipelinedStage(B, dev, depth=2, fine_precision=args.fine_precision)
    pipe.load_constants(host)
    for s_ in pipe.stages:
        s_.instance_ids.copy_(torch.tensor(ids, dtype=torch.float32))

    def post(s_):
        if world > 1:  # NCCL gather of the records on the replica's stream
            return pkg.parallel.gather_records(s_.records, world * B, world)[rank * B:(rank + 1) * B]
        return s_.records

    for _ in range(max(W, 3)):
        pipe.submit(pin, post)
    pipe.drain()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        pipe.submit(pin, post)
    pipe.drain()
    barrier()
    e2e_s = time.perf_counter() - t0
    clock_info = clocks.stop()

    # ---- max over ranks -----------------------------------------------------------------
    tt = torch.tensor([dev_ms, e2e_s * 1e3, wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, wall_ms = tt.tolist()
    if rank != 0:
        return
    total = world * B * K
    value = total / (dev_ms / 1e3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(W, 3),
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "e2e": {"value": total / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / K},
        "gpu_launches": stage.kernels_per_step * K,
        "launches_per_step": stage.kernels_per_step,
        "wall_ms_per_step_incl_flush": wall_ms / K,
        "clocks": clock_info,
    }
    line["roofline"] = dominant_kernel_roofline(pkg, stage, flush)
    line["extra"] = scoring_sweep_point(pkg, stage, dev)
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_arm(args, 2, 2)
        v = r["instances"] / min(r["seconds"])
        line["cpu_baseline"] = {
            "value": v, "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
            "sample": "2 instances through the whole stage, best of 2 (reference sampler needs "
                      "3.5 GB/instance, hence the small sample)"}
    print(json.dumps(line), flush=True)


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured (MEASURED_PEAKS.json)"
    except OSError:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, \
            "fallback (B200_PROFILING.md)"


def dominant_kernel_roofline(pkg, stage, flush):
    """The stage's dominant kernel: fine_tc_kernel (fine_tc.cu), the tcgen05 GEMM (one fp16 product per
    element by default; split-bf16 = three products with --fine-precision split_bf16) with a
    reducing epilogue that pem_fine_fused launches for its two full passes per step (softmax
    statistics; foreground labels + weighted sums) and once more, ragged, for the background-column
    correction.  Each of the two full launches is timed alone with CUDA events on its launch stream
    through the profiling hook pem_fine_fused_pass_ex -- L2 flushed, then the operand preparation kernel
    (untimed), then the launch, i.e. the cache state the kernel meets inside the step; the figure is
    their mean (the cold-L2 durations are reported next to it).

    Algorithmic flops per launch = 2*N1*N2*C per instance (the cosine matrix once; DESIGN.md
    section 4).  In split-bf16 mode the kernel EXECUTES 3x that on the tensor cores (hi*hi + hi*lo +
    lo*hi), reported as executed_mma_tflops next to the measured cuBLAS bf16 peak."""
    L, st = pkg._lib.lib(), stage.stream.cuda_stream
    B, N, C, Nm = stage.B, stage.N, stage.C, stage.Nm
    p = lambda t: t.data_ptr()
    peaks, which = _peaks()

    def fused():
        pkg._lib.check(L.pem_fine_fused_ex(p(stage.feats[0]), p(stage.feats[1]), p(stage.pts[0]), p(stage.pts[1]),
                                           p(stage.model), B, N, N, C, Nm, stage.temp, stage.dis_thres,
                                           stage.fine_precision, p(stage.R), p(stage.t), p(stage.score), None,
                                           None, None, None, p(stage.ws), stage.ws_bytes, st), "fine_fused")

    def one_pass(k):
        pkg._lib.check(L.pem_fine_fused_pass_ex(p(stage.feats[0]), p(stage.feats[1]), B, N, N, C, stage.temp, k,
                                                stage.fine_precision, p(stage.ws), stage.ws_bytes, st),
                       "fine_fused_pass")

    def timed(fn, iters=10, pre=None):
        """mean duration of fn: L2 flushed, then `pre` (untimed; the producer of fn's inputs in the step), then fn"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms = 0.0
        for _ in range(3):
            fn()
        for _ in range(iters):
            flush.zero_()
            if pre is not None:
                pre()
            e0.record(stage.stream)
            fn()
            e1.record(stage.stream)
            stage.stream.synchronize()
            ms += e0.elapsed_time(e1)
        return ms / iters

    with torch.cuda.stream(stage.stream):
        fused_ms = timed(fused)  # also leaves operands and constants in the workspace for the passes
        # as inside the step: the operand preparation has just run (operands L2-resident, as its consumer
        # finds them), everything else flushed
        pass_ms = [timed(lambda k=k: one_pass(k), pre=lambda: one_pass(0)) for k in (1, 4)]
        cold_ms = [timed(lambda k=k: one_pass(k)) for k in (1, 4)]  # operands from HBM
    ms = sum(pass_ms) / 2
    flops = 2.0 * N * N * C * B
    achieved = flops / (ms / 1e3) / 1e12
    peak = float(peaks["bf16_tflops"])
    traffic = None
    single = stage.fine_precision == 1
    nprod = 1.0 if single else 3.0
    summary = "r1_fine_tc_fp16_ncu_summary.json" if single else "r1_fine_tc_ncu_summary.json"
    try:  # DRAM bytes per launch from the committed ncu --set full capture of this kernel
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", summary)) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    return {"kernel": "fine_tc_kernel (tcgen05 cta_group::2 %s GEMM %dx%dx%d per instance, reducing "
                      "epilogue; mean of the stats and the label+wsum launch)"
                      % ("fp16" if single else "split-bf16", N, N, C),
            "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "traffic": traffic, "ms_per_launch": ms,
            "ms_per_launch_by_pass": {"stats": pass_ms[0], "label+wsum": pass_ms[1]},
            "ms_per_launch_by_pass_cold_l2": {"stats": cold_ms[0], "label+wsum": cold_ms[1]},
            "timing": "CUDA events on the launch stream around each launch; before it the L2 is flushed and the "
                      "operand preparation kernel re-run, as in the step (cold_l2: without that)",
            "executed_mma_tflops": nprod * achieved, "executed_frac": nprod * achieved / peak,
            "fused_fine_op_ms": fused_ms,
            "peak_source": which + ", burst bf16 (kernel timed alone)"}


def scoring_sweep_point(pkg, stage, dev):
    """BASELINE configs[3], one point: all 6000 proposals scored (P2 = P1) against 2048 model points."""
    B, P1, M = stage.B, stage.P1, stage.M
    model2 = pkg.synth.make_batch([0], n_model=2048)["model"].to(dev).repeat(B, 1, 1).contiguous()
    mu = pkg.model_utils
    R, t = stage.hyp_R.view(B, P1, 3, 3), stage.hyp_t
    fn = lambda: mu.score_hypotheses(R, t, stage.sparse_pts[0], stage.w1c, model2)
    for _ in range(2):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / 3
    hyps = B * P1
    nfg = float(stage.w1c.sum().item()) / B
    return {"hypotheses_scored_per_sec": hyps / (ms / 1e3), "sweep_point": "%d instances x %d hypotheses x 2048 "
            "model points" % (B, P1), "ms": ms,
            "fp32_pair_evals_per_sec": hyps * nfg * 2048 / (ms / 1e3), "mean_foreground_points": nfg}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="instances per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fine-precision", default="fp16", choices=["fp16", "split_bf16"],
                    help="arithmetic of the fused fine GEMM (include/pem_b200.h PEM_PREC_*); both pass the "
                         "pose parity bars of tests/test_tensorcore_gpu.py")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
