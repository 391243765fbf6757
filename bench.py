#!/usr/bin/env python
"""bench.py -- PEM geometric-stage throughput on B200 (contract: see the task brief).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B]

A *step* is one pass of the geometric stage (openvino-sam-6d_b200/stage.py) over one batch
of B synthetic instances per GPU (default B = 16, BASELINE.json configs[1]: coarse 196 / fine
2048 points, 6000 -> 300 hypotheses, 1024 model points).  Prints ONE JSON line on rank 0.

  value        instances/s, whole job, inputs resident in HBM: K steps round-robin over --in-flight (4) batches
               in flight (one stage replica, streams and CUDA graph per batch; the latency-bound coarse chain
               of one batch runs under the bulk kernels of the others), one CUDA-event pair around the K
               steps, max over ranks.  No L2 flush inside that region: the batches in flight are distinct
               and far larger than L2 (config.l2).  single_batch = the same step with ONE batch in flight
               (the step's latency; L2 flushed between steps)
  e2e          the same through the host-facing call: pinned host inputs -> H2D -> stage ->
               D2H of the 64-byte pose records, all inside the timed region
  roofline     the dominant kernel (the fine-stage tcgen05 GEMM) timed with CUDA events on its launch
               stream; `frac` credits the algorithmic flops ONCE per instance over ALL of its launches in a
               step, `frac_per_launch` credits them once per full launch
  cpu_baseline the reference's CPU arithmetic (the reference's own compiled pointnet2 extension, oracle/_ref, +
               the pinned torch-CPU restatement of model_utils.py) on a bounded sample, host cores stated
  extra        kernels[] (every kernel of the step alone: us, bound, achieved, frac), sustained (the step
               looped >= 2 s, clocks recorded), fp32_peak (measured here), configs3 (hypothesis-scoring
               sweep, BASELINE configs[3]), configs4 (pointnet2 microbench, BASELINE configs[4], with the
               reference's own CUDA kernels compiled for sm_100 timed beside ours), parity_in_bench
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
CPU_KIND_NOTE = ("reference pointnet2 _ext (oracle/_ref, compiled unmodified) + pinned torch-CPU port of "
                 "utils/model_utils.py (bit-identical to the live reference, tests/test_oracle_cpu.py)")


def dtype_label(fine_precision):
    if fine_precision == "fp16":
        return "fp16 GEMM operands, fp32 accumulate (fine similarity); fp32 elsewhere (tf32x2-split tensor-core " \
               "distance products in scoring / PE filter, fp64 3x3 SVD)"
    return "split-bf16 (3 products) GEMM operands, fp32 accumulate (fine similarity); fp32 elsewhere"


def workload_config(args, world):
    return {
        "workload": "synthetic PEM geometric stage, configs[1]: %d instances/GPU, coarse 196 / fine "
                    "2048 points, C=256, 6000->300 hypotheses, 1024 model points" % args.batch,
        "instances_per_gpu": args.batch, "n_gpus": world, "n_dense": 2048, "n_sparse": 196,
        "feat_dim": 256, "n_model": 1024, "n_proposal1": 6000, "n_proposal2": 300,
        "parallelism": "instance-sharded dp%d, pose-record gather" % world,
        "l2": "no flush inside the timed region: %d batches in flight on distinct instances, each step reads 70 MB "
              "of inputs and writes 150 MB of PE groups (working set %d x ~330 MB >> 126 MB L2); "
              "single_batch: flushed between steps" % (getattr(args, "in_flight", 4), getattr(args, "in_flight", 4)),
        "batches_in_flight": getattr(args, "in_flight", 4),
        "fine_gemm": getattr(args, "fine_precision", "fp16") + " tcgen05 (fp32 accumulation; pose parity bars "
                     "0.1 deg / 1e-4 asserted for this mode on the bench batch, tests/test_bench_batch_gpu.py)",
    }


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=100):
        self.index, self.rows, self.proc, self.period = index, [], None, period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", str(self.period)],
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
        pw = []
        for r in self.rows:
            try:
                pw.append(float(r[3]))
            except (ValueError, IndexError):
                pass
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 8 and r[4 + k] == "Active" for r in self.rows)]
        # the busiest half of the samples = "under load"
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_mhz_min": sm[0] if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------- CPU arm
def _oracle_path():
    odir = os.path.join(ROOT, "oracle")
    if odir not in sys.path:
        sys.path.insert(0, odir)


def cpu_arm(args, n_instances, reps):
    """Time the reference CPU arithmetic on `n_instances` instances, `reps` times."""
    _oracle_path()
    import cpu_stage
    pkg = entry.load_package()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = pkg.synth.make_batch(list(range(n_instances)))
    ops = cpu_stage.pointnet2_ops()
    times, out = [], None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = cpu_stage.run_stage_cpu(batch, faithful=True, ops=ops)
        times.append(time.perf_counter() - t0)
    kind = "reference" if ops.kind == "reference" else "port"
    return {"seconds": times, "instances": n_instances, "cores": cores, "kind": kind,
            "kind_detail": CPU_KIND_NOTE if ops.kind == "reference" else "C + torch-CPU port (oracle/)", "result": out}


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
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                         "kind_detail": r["kind_detail"], "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "ranks_running_the_reference": 1,  # at N > 1 this is ONE host process (rank 0), not N
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- GPU arm
def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist
    pkg = entry.load_package()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    numa_cores = pkg.parallel.bind_to_gpu_numa(local_rank) if world > 1 else None
    B, K, W, D = args.batch, args.steps, max(args.warmup, 3), max(1, args.in_flight)
    # D batches in flight: one stage replica (input slots, streams, CUDA graph) per batch, distinct instances.
    # Replica 0 holds this rank's ids rank*B*D .. (+B): the instances the CPU leg and parity_in_bench look at.
    stages, gathers, ids0 = [], [], None
    for d in range(D):
        ids = [(rank * D + d) * B + i for i in range(B)]
        st_ = pkg.stage.GeometricStage(B, dev, fine_precision=args.fine_precision)
        st_.instance_ids.copy_(torch.tensor(ids, dtype=torch.float32))
        st_.load(pkg.synth.make_batch(ids))
        stages.append(st_)
        gathers.append(pkg.parallel.RecordGather(st_, rank, world, mode=args.record_gather))
        if d == 0:
            ids0 = ids
    stage, gather, ids = stages[0], gathers[0], ids0
    host = pkg.synth.make_batch(ids)

    def step_collective(s_=stage, g_=gather):
        # the only multi-GPU traffic: the 64-byte pose records of every rank into one table
        if world > 1:
            with torch.cuda.stream(s_.stream):
                g_.run()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    cur = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def pipelined(n_steps):
        """n_steps steps round-robin over the D replicas (each on its own streams); one event pair on the
        current stream around all of them.  Returns device ms."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        for s_ in stages:
            s_.stream.wait_event(e0)
        for k in range(n_steps):
            s_ = stages[k % D]
            s_.run()
            step_collective(s_, gathers[k % D])
        for s_ in stages:
            cur.wait_stream(s_.stream)
        e1.record(cur)
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1)

    # ---- value: inputs resident, D batches in flight, one CUDA-event pair around the K steps ------------
    clocks = ClockSampler(local_rank)
    clocks.start()
    pipelined(W * D)
    barrier()
    wall0 = time.perf_counter()
    dev_ms = pipelined(K)
    barrier()
    wall = time.perf_counter() - wall0
    records_dev = stage.records.clone()

    # ---- the same step with ONE batch in flight (its latency): L2 flushed between steps, per-step events ----
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    for k in range(K):
        with torch.cuda.stream(stage.stream):
            flush.zero_()
            ev[k][0].record(stage.stream)
        stage.run()
        step_collective()
        ev[k][1].record(stage.stream)
    barrier()
    single_ms = sum(a.elapsed_time(b) for a, b in ev)

    # ---- e2e: pinned host -> H2D -> stage -> D2H records, all inside the timed region --------
    # through the host-facing PipelinedStage (two replicas: copies of batch k+1 overlap batch k)
    upload = args.e2e_upload if args.fine_precision == "fp16" else "fp32"
    pipe = pkg.stage.PipelinedStage(B, dev, depth=2, fine_precision=args.fine_precision, upload=upload)
    pin = pipe.pin_inputs(host)
    h2d = sum(v.numel() * v.element_size() for v in pin.values())
    d2h = B * 16 * 4
    pipe.load_constants(host)
    for s_ in pipe.stages:
        s_.instance_ids.copy_(torch.tensor(ids, dtype=torch.float32))
    pgather = [pkg.parallel.RecordGather(s_, rank, world, mode=args.record_gather) for s_ in pipe.stages]

    def post(s_):
        if world > 1:  # gather of the records on the replica's stream; the host reads its own slice
            g_ = pgather[pipe.stages.index(s_)]
            g_.run()
            return g_.table[rank * B:(rank + 1) * B]
        return s_.records

    for _ in range(W):
        pipe.submit(pin, post)
    last = pipe.drain()
    e2e_records = last.clone()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        pipe.submit(pin, post)
    pipe.drain()
    barrier()
    e2e_s = time.perf_counter() - t0
    clock_info = clocks.stop()

    # ---- sustained: the same step (flush + graph replay + collective) looped for >= sustain seconds -------
    # (every rank must run the SAME number of steps: with the in-graph record gather a step waits for its peers)
    est = torch.tensor([dev_ms / K], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(est, op=dist.ReduceOp.MAX)
    sustained = sustained_run(args, pipelined, barrier, local_rank, est.item())

    # ---- max over ranks -----------------------------------------------------------------
    tt = torch.tensor([dev_ms, e2e_s * 1e3, wall * 1e3, sustained["ms_per_step"], single_ms], dtype=torch.float64,
                      device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, wall_ms, sus_ms, single_ms = tt.tolist()
    # ---- BASELINE configs[3]: hypothesis-scoring sweep, every rank scores its own 16 instances -----
    sweep = scoring_sweep(pkg, stage, dev, world, barrier) if not args.no_extra else None
    if rank != 0:
        return
    total = world * B * K
    value = total / (dev_ms / 1e3)
    sustained["ms_per_step"] = sus_ms
    sustained["value"] = world * B / (sus_ms / 1e3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": dtype_label(args.fine_precision), "data": "synthetic", "config": workload_config(args, world),
        "e2e": {"value": total / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / K,
                "upload": pipe.upload_note()},
        "gpu_launches": stage.kernels_per_step * K,
        "launches_per_step": stage.kernels_per_step,
        "wall_ms_per_step": wall_ms / K,
        "batches_in_flight": D,
        "single_batch": {"ms_per_step": single_ms / K, "value": total / (single_ms / 1e3),
                         "how": "ONE batch in flight (the step's latency): L2 flushed by a 256 MB memset between "
                                "steps, outside the per-step event pairs"},
        "clocks": clock_info,
        "record_gather": gather.how,
        "host_cores_bound": len(numa_cores) if numa_cores else None,
    }
    fp32 = fp32_peak(pkg, stage, dev)
    line["roofline"] = dominant_kernel_roofline(pkg, stage, flush)
    extra = {"sustained": sustained, "fp32_peak": fp32}
    for g_ in gathers:
        if g_.timed_out():
            raise RuntimeError("a pose-record gather gave up waiting for a peer")
    if not args.no_extra:
        extra["kernels"] = kernel_table(pkg, stage, fp32, line["roofline"])
        extra["configs3_scoring_sweep"] = sweep
        extra["configs4_pointnet2_microbench"] = pointnet2_microbench(pkg, dev, fp32)
        extra["hypotheses_scored_per_sec"] = sweep["points"][0]["hypotheses_scored_per_sec"]
    line["extra"] = extra
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_arm(args, 2, 2)
        v = r["instances"] / min(r["seconds"])
        line["cpu_baseline"] = {
            "value": v, "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "kind_detail": r["kind_detail"],
            "sample": "2 instances through the whole stage, best of 2 (reference sampler needs "
                      "3.5 GB/instance, hence the small sample)"}
        cpu_e2e = r["result"]
        if upload == "fp16":  # the host-facing run consumed half-precision features: the reference on the same values
            _oracle_path()
            import cpu_stage
            hb = pkg.stage.PipelinedStage.as_uploaded(pkg.synth.make_batch([0, 1]), "fp16")
            cpu_e2e = cpu_stage.run_stage_cpu(hb, faithful=False)
        extra["parity_in_bench"] = parity_in_bench(r["result"], records_dev.cpu(), cpu_e2e, e2e_records, upload)
    print(json.dumps(line), flush=True)


def parity_in_bench(cpu_result, records, cpu_result_e2e, e2e_records, upload):
    """The instances the CPU leg computed (ids 0..n-1 = the first rows of rank 0's bench batch) against the
    records of the timed GPU steps (device-resident run AND the host-facing pipelined run; the latter against
    the reference fed the features as uploaded)."""
    _oracle_path()
    import pem_oracle as O
    n = cpu_result[0].shape[0]
    out = {"instances": n, "compared_with": "oracle/cpu_stage.run_stage_cpu on the same ids (e2e_run: on the "
                                            "features as uploaded, %s)" % upload}
    for name, rec, cres in (("value_run", records, cpu_result), ("e2e_run", e2e_records, cpu_result_e2e)):
        oR, ot, os_, ost = cres
        rec = rec[:n].float().cpu()
        R, t, s, pid = rec[:, :9].reshape(n, 3, 3), rec[:, 9:12], rec[:, 12], rec[:, 13].long()
        out[name] = {
            "max_dR_deg": O.rotation_angle_deg(R, oR).max().item(),
            "max_dt": (t - ot).abs().max().item(),
            "max_dscore": (s - os_).abs().max().item(),
            "winner_id_match": "%d/%d" % (int((pid == ost["proposal_id"]).sum()), n),
        }
        out[name]["within_north_star_bars"] = bool(out[name]["max_dR_deg"] <= 0.1 and out[name]["max_dt"] <= 1e-4)
    return out


def sustained_run(args, pipelined, barrier, local_rank, ms_est):
    """The timed loop (same batches in flight) run for >= --sustain-seconds of device time: the figure at the
    clocks the GPU HOLDS under this load (the K-step `value` is a burst of ~0.1 s at full boost)."""
    barrier()
    clocks = ClockSampler(local_rank, period_ms=50)
    clocks.start()
    n = max(200, int(args.sustain_seconds * 1e3 / max(ms_est, 0.05)))
    total_ms = pipelined(n)
    barrier()
    info = clocks.stop()
    return {"steps": n, "seconds": total_ms / 1e3, "ms_per_step": total_ms / n, "clocks": info,
            "how": "the timed loop of `value` (graph replays round-robin over the batches in flight), one event "
                   "pair around all steps"}


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured (MEASURED_PEAKS.json)"
    except OSError:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, \
            "fallback (B200_PROFILING.md)"


def _timed(stream, fn, iters=10, pre=None, flush=None, inner=1):
    """mean ms of fn on `stream`: [flush,] [pre (untimed),] event, fn x inner, event.  inner > 1: back-to-back
    launches of an idempotent kernel inside one event pair (an event pair around ONE short launch reads 5-8 us
    whatever the kernel does)."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = 0.0
    with torch.cuda.stream(stream):
        for _ in range(3):
            fn()
        for _ in range(iters):
            if flush is not None:
                flush.zero_()
            if pre is not None:
                pre()
            e0.record(stream)
            for _ in range(inner):
                fn()
            e1.record(stream)
            stream.synchronize()
            ms += e0.elapsed_time(e1)
    return ms / (iters * inner)


def fp32_peak(pkg, stage, dev):
    """FP32 issue peaks measured on this GPU: fused multiply-adds, unfused mul/add (what the
    bit-exact pointnet2 arithmetic is allowed to use), packed FFMA2."""
    L = pkg._lib.lib()
    sink = torch.zeros(4, device=dev)
    ctas, iters = stage.n_sms * 4, 4096
    lane_instr = ctas * 512 * iters * 64
    out = {"how": "pem_bench_fp32_peak: %d CTAs x 512 threads x %d x 64 dependent-chain instructions over 8 "
                  "independent chains per lane, CUDA events, best of 5" % (ctas, iters)}
    for mode, name, flop in ((0, "ffma_tflops", 2), (1, "unfused_tops", 1), (2, "ffma2_tflops", 4)):
        best = 1e9
        for _ in range(5):
            ms = _timed(stage.stream, lambda: pkg._lib.check(
                L.pem_bench_fp32_peak(mode, ctas, iters, sink.data_ptr(), stage.stream.cuda_stream), "fp32 peak"),
                iters=1)
            best = min(best, ms)
        out[name] = lane_instr * flop / (best / 1e3) / 1e12
    out["nominal_ffma_tflops_at_max_clock"] = stage.n_sms * 128 * 2 * 1.965e9 / 1e12
    return out


def dominant_kernel_roofline(pkg, stage, flush):
    """The stage's dominant kernel: fine_tc_kernel (fine_tc.cu), the tcgen05 GEMM (one fp16 product per
    element by default; split-bf16 = three products with --fine-precision split_bf16) with a reducing
    epilogue.  pem_fine_fused launches it THREE times per step: two full passes over the 2048 x 2048 product
    (softmax statistics; foreground labels + weighted sums) and a ragged one over the background columns.
    Each launch is timed alone with CUDA events on its launch stream through the profiling hook
    pem_fine_fused_pass_ex -- L2 flushed, then the operand preparation kernel (untimed), then the launch, i.e.
    the cache state the kernel meets inside the step.

    SURVEY 8d: "2*(M+1)^2*C flop algorithmic per instance; a two-pass implementation executes 2x but is
    credited 1x".  So `achieved` = 2*N1*N2*C*B / (sum of the three launch durations) and `frac` = that over the
    measured bf16 peak; `frac_per_launch` credits the same flops to each full launch on its own (how busy the
    tensor pipe is while the kernel runs).  In split-bf16 mode the kernel EXECUTES 3x the products."""
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

    fused_ms = _timed(stage.stream, fused, flush=flush)  # also leaves operands and constants in the workspace
    prep = lambda: one_pass(0)
    names = {1: "stats", 4: "label+wsum", 5: "ragged_bg_correction"}
    pass_ms = {names[k]: _timed(stage.stream, lambda k=k: one_pass(k), pre=prep, flush=flush) for k in (1, 4, 5)}
    cold_ms = {names[k]: _timed(stage.stream, lambda k=k: one_pass(k), flush=flush) for k in (1, 4)}
    prep_ms = _timed(stage.stream, prep, flush=flush)
    flops = 2.0 * N * N * C * B
    total_ms = sum(pass_ms.values())
    achieved = flops / (total_ms / 1e3) / 1e12
    per_launch = flops / ((pass_ms["stats"] + pass_ms["label+wsum"]) / 2 / 1e3) / 1e12
    peak = float(peaks["bf16_tflops"])
    single = stage.fine_precision == 1
    nprod = 1.0 if single else 3.0
    traffic, traffic_src = None, None
    for name in ("r2b_fine_tc_fp16_ncu_summary.json", "r2_fine_tc_fp16_ncu_summary.json", "r1_fine_tc_fp16_ncu_summary.json") if single else \
            ("r2_fine_tc_ncu_summary.json", "r1_fine_tc_ncu_summary.json"):
        try:  # DRAM bytes per launch from the committed ncu --set full capture of this kernel
            with open(os.path.join(ROOT, "profiles", name)) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
            traffic_src = "profiles/" + name + " (ncu --set full capture, not measured in this run)"
            break
        except (OSError, ValueError):
            pass
    return {"kernel": "fine_tc_kernel (tcgen05 cta_group::2 %s GEMM %dx%dx%d per instance, reducing "
                      "epilogue; 3 launches per step: stats, label+wsum, ragged background-column correction)"
                      % ("fp16" if single else "split-bf16", N, N, C),
            "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
            "credit": "algorithmic 2*N1*N2*C flop per instance ONCE over all launches of the kernel in a step "
                      "(SURVEY 8d)",
            "algorithmic_gflop_per_step": flops / 1e9,
            "ms_all_launches": total_ms, "ms_per_launch_by_pass": pass_ms,
            "frac_per_launch": per_launch / peak, "achieved_per_launch": per_launch,
            "ms_per_launch_by_pass_cold_l2": cold_ms,
            "timing": "CUDA events on the launch stream around each launch; before it the L2 is flushed and the "
                      "operand preparation kernel re-run, as in the step (cold_l2: without that)",
            "executed_mma_tflops": nprod * 2 * flops / ((pass_ms["stats"] + pass_ms["label+wsum"]) / 1e3) / 1e12,
            "executed_frac": nprod * 2 * flops / ((pass_ms["stats"] + pass_ms["label+wsum"]) / 1e3) / 1e12 / peak,
            "fused_fine_op_ms": fused_ms, "operand_prep_ms": prep_ms,
            "peak_source": which + ", burst bf16 (kernel timed alone)"}


# ------------------------------------------------------------------------------- per-kernel table
def kernel_table(pkg, stage, fp32, roof):
    """Every launch of the step timed ALONE with CUDA events on the stage's stream (10 back-to-back launches per
    event pair, inputs as the step left them, L2-warm like inside the step), with the bound and algorithmic work
    of SURVEY 8d.  The fused fine
    call is split into its kernels through the pass hook.  The step itself is shorter than the sum: the
    side-stream kernels overlap the coarse chain."""
    B, N, M, C, Nm, P1, P2 = stage.B, stage.N, stage.M, stage.C, stage.Nm, stage.P1, stage.P2
    peaks, _ = _peaks()
    hbm = float(peaks["hbm_gbs"])
    ffma, unf = fp32["ffma_tflops"], fp32["unfused_tops"]
    st = stage.stream.cuda_stream
    stage.run()
    stage.stream.synchronize()
    nfg = float(stage.w1c.sum().item()) / B
    rows = []

    def row(name, us, bound, work=None, unit=None, peak=None, note=None):
        r = {"kernel": name, "us": us, "bound": bound}
        if work is not None:
            r["achieved"] = work / (us * 1e-6) / (1e12 if unit != "GB/s" else 1e9)
            r["unit"], r["peak"], r["frac"] = unit, peak, r["achieved"] / peak
        if note:
            r["note"] = note
        rows.append(r)

    for name, fn, k in stage.launches():
        if name == "fine_fused":
            continue
        if name.startswith("gather_poses"):
            # N > 1 GPUs: the record push / flag wait pair only means something inside a step that every rank runs --
            # alone, the wait kernel sits out its 2 s time-out on every launch (r2b_scale_n2.json: "us": 2000008)
            continue
        us = 1e3 * _timed(stage.stream, lambda fn=fn, name=name: pkg._lib.check(fn(st), name), iters=5, inner=10)
        if name.startswith("fps"):
            row(name, us, "fp32 (unfused, latency: M-1 dependent steps)", 2 * B * (M - 1) * N * 10.0, "Top/s", unf,
                "streaming-equivalent %.0f GB/s" % (2 * B * (M - 1) * N * 20.0 / (us * 1e-6) / 1e9))
        elif name.startswith("gather_rows"):
            row(name, us, "hbm", 2 * B * (4.0 * M + 8.0 * C * M), "GB/s", hbm)
        elif name == "coarse_similarity":
            row(name, us, "fp32", 2.0 * B * (M + 1) ** 2 * C, "TFLOP/s", ffma, "2 launches: row norms + FFMA GEMM")
        elif name == "coarse_hypotheses":
            row(name, us, "latency/ALU", note="%.1f M proposals/s" % (B * P1 / us))
        elif name == "score_hypotheses":
            row(name, us, "fp32 (executed on the tensor pipe as a split-tf32 distance GEMM)",
                B * P2 * (196.0 * Nm * 8 + 196 * 15), "TFLOP/s", ffma,
                "%.1f M hypotheses/s; %.0f foreground points of 196 on average" % (B * P2 / us, nfg))
        elif name.startswith("pe_group2_masked"):
            row(name, us, "fp32 re-test + hbm output", B * (4.0 * N * 96 + 8.0 * 3 * N * 96 * 2), "GB/s", hbm,
                "both PE scales of one cloud set: idx + (B,6,N,32|64) outputs; N^2 scan replaced by the mask filter")
        elif name == "pe_masks":
            row(name, us, "tensor (split-tf32) + mask building", 2.0 * 2 * B * N * N * 8 * 3, "TFLOP/s",
                float(peaks["bf16_tflops"]) / 2, "candidate filter for 2B clouds; tf32 peak taken as bf16/2")
        else:
            row(name, us, "latency")
    # the fused fine call, kernel by kernel
    pm = roof["ms_per_launch_by_pass"]
    flops = 2.0 * N * N * C * B
    row("tc_prep_kernel (normalise + fp16 convert + bg dots)", 1e3 * roof["operand_prep_ms"], "hbm",
        2.0 * B * (N + 1) * C * 4 + 2.0 * B * N * C * (2 if stage.fine_precision == 1 else 4), "GB/s", hbm)
    row("fine_tc_kernel<stats>", 1e3 * pm["stats"], "tensor", flops, "TFLOP/s", float(peaks["bf16_tflops"]),
        "credited the full product on its own launch (frac_per_launch)")
    row("fine_tc_kernel<label+wsum>", 1e3 * pm["label+wsum"], "tensor", flops, "TFLOP/s",
        float(peaks["bf16_tflops"]), "credited the full product on its own launch (frac_per_launch)")
    row("fine_tc_kernel<wsum, ragged>", 1e3 * pm["ragged_bg_correction"], "tensor",
        note="background columns only (one 256-column strip per instance as a rule)")
    rest = 1e3 * (roof["fused_fine_op_ms"] - roof["operand_prep_ms"] - sum(pm.values()))
    row("tc_finish_* x3 + tc_bg_gather + weighted_procrustes + fine_pose_score", rest, "latency / fp32",
        B * N * (30.0 + 8.0 * Nm), "TFLOP/s", ffma, "remainder of the fused fine call (6 small kernels)")
    return rows


# ------------------------------------------------------------------------------- configs[3]
def scoring_sweep(pkg, stage, dev, world, barrier):
    """BASELINE configs[3]: hypotheses per instance in {6k,...,60k}, ALL scored (P2 = P1) against 2048 model
    points, 16 instances per GPU, argmax on the device.  Hypotheses come from the stage's own sampler +
    Procrustes on the bench batch.  Aggregate over ranks = sum of hypotheses / max time."""
    import torch.distributed as dist
    L = pkg._lib.lib()
    B, M = stage.B, stage.M
    st = stage.stream.cuda_stream
    p = lambda t: t.data_ptr()
    model2 = pkg.synth.make_batch([0], n_model=2048)["model"].to(dev).repeat(B, 1, 1).contiguous()
    g = torch.Generator(device="cpu").manual_seed(3)
    points = []
    for H in (6000, 12000, 24000, 36000, 48000, 60000):
        rand = torch.rand(B, 3 * H, generator=g).to(dev)
        R, t = torch.empty(B, H, 9, device=dev), torch.empty(B, H, 3, device=dev)
        resid, scores = torch.empty(B, H, device=dev), torch.empty(B, H, device=dev)
        bR, bt = torch.empty(B, 9, device=dev), torch.empty(B, 3, device=dev)
        pid = torch.empty(B, dtype=torch.int32, device=dev)
        sel = torch.arange(H, dtype=torch.int32, device=dev).repeat(B, 1).contiguous()
        with torch.cuda.stream(stage.stream):
            pkg._lib.check(L.pem_coarse_hypotheses(p(stage.cdf), p(rand), p(stage.sparse_pts[0]),
                                                   p(stage.sparse_pts[1]), B, M, M, H, p(R), p(t), p(resid), None, st),
                           "hypotheses")

        def fn():
            pkg._lib.check(L.pem_score_hypotheses(p(R), p(t), None, B, H, H, p(stage.sparse_pts[0]), p(stage.w1c), M,
                                                  p(model2), 2048, p(scores), st), "score")
            pkg._lib.check(L.pem_select_best(p(scores), p(sel), p(R), p(t), B, H, H, p(bR), p(bt), p(pid), st), "best")

        barrier()
        ms = _timed(stage.stream, fn, iters=3)
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = tt.item()
        points.append({"hypotheses_per_instance": H, "ms": ms,
                       "hypotheses_scored_per_sec": world * B * H / (ms / 1e3),
                       "fp32_tflops_credited": world * B * H * (196.0 * 2048 * 8 + 196 * 15) / (ms / 1e3) / 1e12})
    out = {"config": "B=%d instances/GPU x %d GPU(s), M=196 observed points, 2048 model points, P2 = P1, device "
                     "argmax included" % (B, world), "points": points,
           "credit": "SURVEY 8d: H*196*Nm*8 + H*196*15 flop per hypothesis (3.21 MFLOP at Nm = 2048)"}
    return out


def cpu_scoring_baseline():
    """Reference scoring stage (MU:267-275) on the host cores, bounded sample: 1 instance x 512 hypotheses."""
    _oracle_path()
    import pem_oracle as O
    g = torch.Generator().manual_seed(0)
    H = 512
    R = torch.linalg.qr(torch.randn(1, H, 3, 3, generator=g))[0]
    t = 0.1 * torch.randn(1, H, 3, generator=g)
    pts, model = 0.4 * torch.randn(1, 196, 3, generator=g), 0.4 * torch.randn(1, 2048, 3, generator=g)
    w1 = torch.ones(1, 196)
    O.score_hypotheses(R[:, :64], t[:, :64], pts, w1, model)
    t0 = time.perf_counter()
    O.score_hypotheses(R, t, pts, w1, model)
    return {"hypotheses_scored_per_sec": H / (time.perf_counter() - t0), "cores": os.cpu_count(),
            "sample": "1 instance x 512 hypotheses x 2048 model points, torch CPU (the reference's arithmetic)"}


# ------------------------------------------------------------------------------- configs[4]
def pointnet2_microbench(pkg, dev, fp32):
    """BASELINE configs[4]: FPS / ball_query / grouping / gather on 64 clouds x 16384 points -> 2048 samples.
    Ours through the `pointnet2._ext` mirror; beside it the REFERENCE's own CUDA kernels compiled for sm_100
    (oracle/_ref_cuda, built from the unmodified sources by oracle/build_ref.py) when that build is present;
    indices checked against the reference CPU ops (oracle/_ref) on the first 2 clouds."""
    _oracle_path()
    peaks, _ = _peaks()
    hbm = float(peaks["hbm_gbs"])
    ext = pkg.pointnet2_ext
    Bc, N, M = 64, 16384, 2048
    g = torch.Generator().manual_seed(5)
    u = torch.randn(Bc, N, 3, generator=g)
    u = u / u.norm(dim=2, keepdim=True)
    axes = 0.3 + 0.7 * torch.rand(Bc, 1, 3, generator=g)
    xyz_h = (u * axes)
    xyz_h = (xyz_h / xyz_h.norm(dim=2).amax(1).view(Bc, 1, 1)).contiguous()   # radius-normalised
    xyz = xyz_h.to(dev)
    feat3 = xyz.transpose(1, 2).contiguous()
    feat256 = torch.randn(Bc, 256, N, generator=g).to(dev)
    ref = None
    try:
        import build_ref
        if os.path.exists(build_ref.cuda_so_path()):
            ref = build_ref.load_ref_cuda_ext()
    except Exception as e:  # pragma: no cover
        ref = None
        ref_err = repr(e)
    cur = torch.cuda.current_stream(dev)

    def t_ms(fn, iters=5):
        return _timed(cur, fn, iters=iters)

    rows = []

    def add(name, ours, theirs, bound, work, unit, peak, equal=None, note=None):
        r = {"op": name, "ms": t_ms(ours), "bound": bound}
        r["achieved"] = work / (r["ms"] / 1e3) / (1e9 if unit == "GB/s" else 1e12)
        r["unit"], r["peak"], r["frac"] = unit, peak, r["achieved"] / peak
        if theirs is not None and ref is not None:
            r["reference_cuda_sm100_ms"] = t_ms(theirs)
            r["speedup_vs_reference_cuda"] = r["reference_cuda_sm100_ms"] / r["ms"]
        if equal is not None:
            r["matches_reference_cuda"] = equal
        if note:
            r["note"] = note
        rows.append(r)

    idx = ext.furthest_point_sampling(xyz, M)
    eq = None
    if ref is not None:
        ridx = ref.furthest_point_sampling(xyz, M)
        mine0 = ext.furthest_point_sampling(xyz, M, skip_near_origin=False)
        same = (mine0 == ridx).all(1)
        eq = "%d/%d clouds identical to the reference CUDA kernel (flags=0; it contracts the distance into FMAs " \
             "and breaks ties by thread slot, so near-ties may differ)" % (int(same.sum()), Bc)
    add("furthest_point_sampling 64x16384->2048", lambda: ext.furthest_point_sampling(xyz, M),
        (lambda: ref.furthest_point_sampling(xyz, M)) if ref else None, "fp32 (unfused) / latency",
        Bc * (M - 1) * N * 10.0, "Top/s", fp32["unfused_tops"], eq,
        "streaming-equivalent GB/s = B*(M-1)*N*20 B / t")
    rows[-1]["streaming_equivalent_gbs"] = Bc * (M - 1) * N * 20.0 / (rows[-1]["ms"] / 1e3) / 1e9
    new_xyz = ext.gather_points(feat3, idx).transpose(1, 2).contiguous()
    for r_, ns in ((0.1, 32), (0.2, 64)):
        eq = None
        mine = ext.ball_query(new_xyz, xyz, r_, ns)
        if ref is not None:
            theirs = ref.ball_query(new_xyz, xyz, r_, ns)
            eq = "%.4f %% of neighbour slots identical to the reference CUDA kernel (FMA-contracted distances " \
                 "there: points on the ball surface may differ)" % (100.0 * (mine == theirs).float().mean().item())
        add("ball_query r=%.1f ns=%d (2048 centres, 16384 pts)" % (r_, ns),
            lambda r_=r_, ns=ns: ext.ball_query(new_xyz, xyz, r_, ns),
            (lambda r_=r_, ns=ns: ref.ball_query(new_xyz, xyz, r_, ns)) if ref else None,
            "fp32 (unfused), upper bound without early exit", Bc * M * N * 9.0, "Top/s", fp32["unfused_tops"], eq)
    bq = ext.ball_query(new_xyz, xyz, 0.2, 64)
    for name, f in (("C=3", feat3), ("C=256", feat256)):
        Cc = f.shape[1]
        eq = None
        if ref is not None:
            eq = bool(torch.equal(ext.group_points(f, bq), ref.group_points(f, bq)))
        add("group_points %s ns=64" % name, lambda f=f: ext.group_points(f, bq),
            (lambda f=f: ref.group_points(f, bq)) if ref else None, "hbm",
            Bc * (4.0 * M * 64 + 8.0 * Cc * M * 64), "GB/s", hbm, eq)
        eq = None
        if ref is not None:
            eq = bool(torch.equal(ext.gather_points(f, idx), ref.gather_points(f, idx)))
        add("gather_points %s (B,C,N)->(B,C,2048)" % name, lambda f=f: ext.gather_points(f, idx),
            (lambda f=f: ref.gather_points(f, idx)) if ref else None, "hbm",
            Bc * (4.0 * M + 8.0 * Cc * M), "GB/s", hbm, eq)
    # indices vs the reference CPU ops on a bounded sample (2 clouds)
    cpu_check = None
    try:
        import cpu_stage
        ops = cpu_stage.pointnet2_ops()
        t0 = time.perf_counter()
        cidx = ops.fps(xyz_h[:2], M)
        cpu_fps_s = time.perf_counter() - t0
        cbq = ops.ext.ball_query(new_xyz[:2].cpu(), xyz_h[:2], 0.2, 64) if ops.kind == "reference" else None
        cpu_check = {"fps_indices_bit_exact_vs_reference_cpu": bool(torch.equal(cidx, idx[:2].cpu())),
                     "ball_query_bit_exact_vs_reference_cpu": None if cbq is None else bool(torch.equal(cbq, bq[:2].cpu())),
                     "clouds_checked": 2, "reference_cpu_fps_s_per_cloud": cpu_fps_s / 2, "ops": ops.kind}
    except Exception as e:  # pragma: no cover
        cpu_check = {"error": repr(e)}
    return {"config": "64 clouds x 16384 points -> 2048 samples; ball-query centres = the sampled points",
            "reference_cuda": "oracle/_ref_cuda (the reference's .cu compiled for sm_100, unmodified)" if ref else
                              "not available in this run", "rows": rows, "cpu_check": cpu_check,
            "cpu_scoring_baseline": cpu_scoring_baseline()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="instances per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the per-kernel table, the sweep and the microbench")
    ap.add_argument("--sustain-seconds", type=float, default=2.0)
    ap.add_argument("--in-flight", type=int, default=4,
                    help="batches in flight in the `value` loop (one stage replica, streams and CUDA graph each); "
                         "1 = one batch at a time")
    ap.add_argument("--e2e-upload", default="fp16", choices=["fp16", "fp32"],
                    help="feature upload format of the host-facing path (fp16 only with the fp16 fine GEMM)")
    ap.add_argument("--record-gather", default="peer", choices=["peer", "nccl"],
                    help="N > 1: peer stores + flag wait inside the step's graph, or an NCCL all-gather after it")
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
