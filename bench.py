#!/usr/bin/env python
"""bench.py -- DDP solves/sec on B200 (BASELINE.json metric), one JSON line on stdout.

A "step" is one pass of the hot path over one batch: the node's two-stage IPDDP protocol
(teach_repeat_planner.cpp:853-951) for `--batch` independent 100-knot corridor problems per GPU
(BASELINE.json configs[1]; the reference's differentially-flat 9-state/10-input quadrotor model, SURVEY.md
section 0).  `value` = whole-job solves/s with inputs resident in HBM (device entry point, CUDA events);
`e2e` = the same through the host-buffer C-ABI call (pinned host memory, H2D + solve + D2H in the timed
region).  `--impl reference` times the reference's own CPU implementation (oracle/_ref = its unmodified
translation unit; else the C restatement) on the host cores for the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DDP solves/sec (100-knot quadrotor, batch 4096)"
F_CONST = 4 * 9**3 + 8 * 81 * 10 + 4 * 9 * 100 + 2 * 81 + 8 * 90 + 1000.0 / 3.0 + 2 * 100 * 10  # SURVEY.md 8(d), n=9, m=10
F_ROW = 2 * 100 + 4 * 90 + 2 * 81 + 70 + 63


def bwd_flops_per_knot(planes_per_cell: float) -> float:
    """Algorithmic (dense) flop count of one knot of the backward pass, SURVEY.md section 8(d)."""
    return F_CONST + (6.0 * planes_per_cell + 55.0) * F_ROW


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def host_cores() -> int:
    """Host threads this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU arms pass the count explicitly)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def reference_arm(args):
    """The reference's own CPU implementation of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from direct_b200.problems import STAGE0, STAGE1, make_batch
    from oracle import oracle_py as O
    O.build(ref=True)
    use_ref = O.ref_available()
    cores = host_cores()
    sample = args.ref_sample
    pb = make_batch(sample, args.knots, args.kind)

    def step():
        if use_ref:
            r0 = O.solve_batch(pb, nthreads=cores, use_ref=True, infeas=1, zero_init=1, **STAGE0)
            dur = np.where((r0.rtn == 2)[:, None], r0.poly_time, pb.durations)
            return O.solve_batch(pb, nthreads=cores, use_ref=True, infeas=r0.infeas_out, zero_init=0,
                                 init_bez=r0.bez_coeff, durations=dur, **STAGE1)
        return O.two_stage_batch(pb, nthreads=cores)[1]

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = sample / dt
    # the dense C restatement for context (no per-call allocation: faster than the reference's Eigen code)
    port_sample = max(sample, args.port_sample)   # a throughput figure needs many more trajectories than threads (heavy tail)
    pbp = pb if port_sample == sample else make_batch(port_sample, args.knots, args.kind)
    O.two_stage_batch(pbp.slice(0, min(port_sample, 4 * cores)), nthreads=cores)   # warm-up
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.two_stage_batch(pbp, nthreads=cores)
    port_value = port_sample * args.steps / (time.perf_counter() - t0)
    # The reference's translation unit is compiled here against oracle/shim/Eigen/Dense, an eager stand-in for Eigen3 (the real
    # library is not in this image) that allocates far more than Eigen's expression templates would: it UNDERSTATES the
    # reference's speed.  The allocation-free C restatement of the same algorithm OVERSTATES it.  The arm's value is the
    # faster of the two, so that a ratio computed from it can only be conservative; both numbers are on the line.
    ref_value = value if use_ref else None
    if port_value >= value:
        value, kind, dt, sample = port_value, "port", port_sample / port_value, port_sample
    else:
        kind = "reference"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(args, per_gpu=args.batch), reference_sample_per_step=sample),
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": kind,
                         "sample": f"first {sample} trajectories of the workload per step ({args.ref_sample} for the reference's translation "
                                   f"unit, {port_sample} for the C restatement), {cores} OpenMP threads; value = the faster of "
                                   "(a) the reference's ddp_optimizer.cpp compiled unmodified against oracle/shim (eager Eigen stand-in) "
                                   "and (b) the allocation-free C restatement oracle/ipddp_oracle.c",
                         "reference_tu_value": ref_value, "port_value": port_value,
                         "linear_algebra_stand_in": bool(use_ref),
                         "note": "reference_tu_value is a LOWER bound of the reference's speed (stand-in Eigen), port_value an UPPER "
                                 "bound (no per-call allocation, ddp_optimizer.cpp:479-504 copies absent)"},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(args, per_gpu):
    return {"workload": f"two-stage IPDDP (stage 0 feasibility + stage 1 optimise), {per_gpu} trajectories/GPU x "
                        f"{args.knots} knots, 9-state/10-input flat quadrotor model, {args.kind} corridor "
                        f"(P={'6' if args.kind == 'box' else '6..14'} planes/cell), weights of global_planner.launch",
            "batch_per_gpu": per_gpu, "knots": args.knots, "corridor": args.kind, "precision": args.precision,
            "l2": "256 MiB buffer written between timed steps (L2 flush); the per-warp workspaces (1 GB) exceed L2 anyway"}


def model_b_report(capi, torch, dev, local_rank, flush, steps, with_cpu):
    """SURVEY.md section 8(d) config 2, model (B): BASELINE.json's literal 12-state / 4-input quadrotor, fp32, batch 4096 x
    100 knots, unconstrained DDP (include/direct_gddp.h).  The reference has no such model: NO reference
    parity, checker = oracle/gddp_oracle.c.  Reported next to the headline, never instead of it."""
    import dataclasses
    from direct_b200 import gddp
    B, N = 4096, 100
    # the start / goal distribution BASELINE.md section 3 states (15-25 m transfers), not the 1.5-3.5 m hops of round 1
    gp = dataclasses.replace(gddp.make_quad_batch(B, N, workload="stated"), tol=1e-5)
    solver = capi.Solver(local_rank, "fp32")
    t = {k: torch.from_numpy(getattr(gp, k)).to(dev) for k in ("x0", "xg")}
    o = dict(rtn=torch.zeros(B, dtype=torch.int32, device=dev), iters=torch.zeros(B, dtype=torch.int32, device=dev),
             cost=torch.zeros(B, dtype=torch.float64, device=dev), x=torch.zeros(B, N + 1, 12, dtype=torch.float64, device=dev),
             u=torch.zeros(B, N, 4, dtype=torch.float64, device=dev), stats=torch.zeros(B, 4, dtype=torch.int64, device=dev))
    pc = gddp.problem_struct(gp, t["x0"].data_ptr(), t["xg"].data_ptr(), 0)
    oc = gddp.ResultC(*[o[n].data_ptr() for n, _ in gddp.ResultC._fields_])
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        flush.fill_(1)
        gddp.solve_device(solver, pc, oc, stream)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        flush.fill_(k & 255)
        ev[k][0].record()
        gddp.solve_device(solver, pc, oc, stream)
        ev[k][1].record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    stats = o["stats"].cpu().numpy()
    rtn = o["rtn"].cpu().numpy()
    knots = int(stats[:, 2].sum())
    fl = gddp.bwd_flops_per_knot(12, 4)
    peak = solver.fma_peak_tflops("fp32")
    # e2e: the host-buffer C-ABI call from / into pinned host memory (H2D + solve + D2H inside the call)
    hres = gddp.GddpResult(B, N, 12, 4)
    pins = []
    for name in ("rtn", "iters", "cost", "x", "u", "stats"):
        a = getattr(hres, name)
        tp = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
        pins.append(tp)
        setattr(hres, name, tp.numpy())
    gpin = dataclasses.replace(gp, x0=torch.from_numpy(gp.x0).pin_memory().numpy(), xg=torch.from_numpy(gp.xg).pin_memory().numpy())
    gddp.solve(solver, gpin, hres)         # warm-up (staging buffers of the handle)
    torch.cuda.synchronize()
    e2e_n = max(3, steps)
    t0 = time.perf_counter()
    for _ in range(e2e_n):
        flush.fill_(2)
        gddp.solve(solver, gpin, hres)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_n
    st = solver.stats()
    traffic, traffic_src = None, None
    try:   # dram bytes of gddp_pair_kernel per launch, one ncu --set full capture of this workload (profiles/)
        prof = json.load(open(os.path.join(ROOT, "profiles", "latest_ncu_model_b.json")))
        if prof.get("workload") == f"{B}x{N} quad fp32 stated":
            traffic, traffic_src = prof["dram_bytes_per_launch"], prof["source"]
    except Exception:
        pass
    out = {"workload": f"unconstrained DDP, {B} x {N}-knot 12-state/4-input rigid-body quadrotor (explicit Euler, dt 0.05 -- SURVEY.md 8(d) "
                       "names RK4; the kernel's Jacobian sparsity is Euler's), start at rest in [-10,10]^2 x [0.5,2.5], goal at rest 15-25 m "
                       "away (BASELINE.md section 3), iter_max 50; model (B) of SURVEY.md 8(d): the reference has no such model, parity "
                       "unpinned",
           "dtype": "f32", "value": B / ms * 1e3, "unit": "solves/s", "ms_per_step": ms, "gpu_launches": steps,
           "e2e": {"value": B / e2e_s, "unit": "solves/s", "h2d_bytes_per_step": int(st.h2d_bytes), "d2h_bytes_per_step": int(st.d2h_bytes),
                   "ms_per_step": e2e_s * 1e3},
           "roofline": {"bound": "fp32_fma", "achieved": fl * knots / (ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                        "frac": fl * knots / (ms * 1e-3) / 1e12 / peak if peak else None, "traffic": traffic, "traffic_source": traffic_src,
                        "flops_per_bwd_knot": fl, "bwd_knots_per_launch": knots},
           "solve_stats": {"converged_frac": float((rtn == 1).mean()), "mean_iters": float(o["iters"].float().mean().item()),
                           "bwd_sweeps_per_solve": float(stats[:, 0].mean()), "rollouts_per_solve": float(stats[:, 1].mean())}}
    if with_cpu:
        from oracle import gddp_py as G   # checker / CPU baseline only
        cores = host_cores()
        t0 = time.perf_counter()
        a = G.solve_batch(gp, nthreads=cores)
        cpu_s = time.perf_counter() - t0
        okm = (a.rtn == 1) & (hres.rtn == 1)
        rel = np.abs(hres.cost[okm] - a.cost[okm]) / np.abs(a.cost[okm])
        out["cpu_baseline"] = {"value": B / cpu_s, "unit": "solves/s", "cores": cores, "kind": "port",
                               "sample": f"all {B} problems, fp64 C oracle (oracle/gddp_oracle.c), {cores} OpenMP threads"}
        out["parity_sample"] = {"n": int(okm.sum()), "tol": 1e-3, "frac_within_tol": float((rel <= 1e-3).mean())}
    solver.close()
    return out


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The one JSON line of this process, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # stdout carries exactly one JSON line: whatever libraries print on fd 1 (NCCL's version banner under torchrun) goes to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="trajectories per GPU")
    ap.add_argument("--knots", type=int, default=100)
    ap.add_argument("--kind", default="box", choices=["box", "poly"])
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"])
    ap.add_argument("--cpu-sample", type=int, default=4096)
    ap.add_argument("--ref-sample", type=int, default=48)
    ap.add_argument("--port-sample", type=int, default=1024, help="trajectories per step of the C restatement in --impl reference")
    ap.add_argument("--screen-sample", type=int, default=1024, help="trajectories of the parity sample run through the conditioning screen")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-model-b", action="store_true", help="skip the secondary 12-state quadrotor (model (B)) report")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from direct_b200 import capi
    from direct_b200 import dist as D
    from direct_b200.problems import STAGE0, STAGE1, TIME_POWER, make_batch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- options: rank 0's are broadcast (NCCL) so every shard is solved identically ----------------------
    opts = dict(w_snap0=STAGE0["w_snap"], w_terminal0=STAGE0["w_terminal"], w_time0=STAGE0["w_time"],
                iter_max0=STAGE0["iter_max"], w_snap=STAGE1["w_snap"], w_terminal=STAGE1["w_terminal"],
                w_time=STAGE1["w_time"], iter_max=STAGE1["iter_max"], time_power=TIME_POWER, max_vel=2.0, max_acc=2.0)
    if world > 1:
        opts = D.broadcast_options(opts if rank == 0 else None, dev)
    ts = capi.TwoStage(opts["w_snap0"], opts["w_terminal0"], opts["w_time0"], opts["iter_max0"], opts["w_snap"],
                       opts["w_terminal"], opts["w_time"], opts["iter_max"], opts["time_power"])

    # ---- this rank's shard of the synthetic workload (weak scaling: `batch` trajectories per GPU) ---------
    B, N = args.batch, args.knots
    pb = make_batch(B, N, args.kind, first=rank * B, max_vel=opts["max_vel"], max_acc=opts["max_acc"])
    solver = capi.Solver(local_rank, args.precision)

    # device-resident inputs / outputs (torch owns the memory; the library gets raw pointers)
    d_in = {k: torch.from_numpy(getattr(pb, k)).to(dev) for k in ("planes", "nplanes", "durations", "x0", "xd")}
    # the four gathered result fields live in ONE packed buffer per rank (views), so that the gather is a single collective
    packed = D.PackedResults({"poly_time": ((N,), torch.float64), "bez_coeff": ((N, 18), torch.float64), "rtn": ((), torch.int32),
                              "cost": ((), torch.float64)}, B, B, dev)
    d_out = dict(rtn=packed.view("rtn"), infeas_out=torch.zeros(B, dtype=torch.int32, device=dev),
                 line_failed_out=torch.zeros(B, dtype=torch.int32, device=dev), iters=torch.zeros(B, dtype=torch.int32, device=dev),
                 cost=packed.view("cost"), x_final=torch.zeros(B, 9, dtype=torch.float64, device=dev),
                 poly_coeff=torch.zeros(B, N, 18, dtype=torch.float64, device=dev),
                 bez_coeff=packed.view("bez_coeff"),
                 poly_time=packed.view("poly_time"), jerk=torch.zeros(B, N, dtype=torch.float64, device=dev),
                 stats=torch.zeros(B, 8, dtype=torch.int64, device=dev))
    batch = capi.Batch(B, N, pb.P_max, d_in["planes"].data_ptr(), d_in["nplanes"].data_ptr(), d_in["durations"].data_ptr(),
                       None, d_in["x0"].data_ptr(), d_in["xd"].data_ptr(), None, None, 1, pb.max_vel, pb.max_acc,
                       1.0, 1.0, 1.0, 0, opts["time_power"], 1, 0, 0)
    out1 = capi.ResultC(*[d_out[n].data_ptr() for n, _ in capi.ResultC._fields_])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    counts = [B] * world

    def gather():
        if world > 1:  # gather of the solved trajectories to rank 0 (one NCCL collective into a preallocated buffer), part of the step
            return packed.gather(counts)
        return None

    def device_step():
        stream = torch.cuda.current_stream().cuda_stream
        solver.solve_two_stage_device(batch, ts, None, out1, stream)
        gather()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        flush.fill_(1)
        device_step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    kernel_ms = []
    for k in range(args.steps):
        flush.fill_(k & 255)          # L2 flush between timed iterations (outside the event pair)
        ev[k][0].record()
        device_step()
        ev[k][1].record()
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_s = sum(step_ms) / 1e3
    st = solver.stats()
    kernel_ms = st.kernel_ms            # solve kernel of the last step (library's own events, same stream)
    clocks = sampler.stop()
    if world > 1:
        total_s = D.max_over_ranks(total_s, dev)
    value = world * B * args.steps / total_s

    # ---- e2e: the host-buffer C-ABI call (what the reference-facing shim uses), pinned host memory --------
    def pinned_like(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t
    keep = {k: pinned_like(getattr(pb, k)) for k in ("planes", "nplanes", "durations", "seeds", "x0", "xd")}
    from direct_b200.problems import ProblemBatch
    pbp = ProblemBatch(B, N, pb.P_max, *[keep[k].numpy() for k in ("planes", "nplanes", "durations", "seeds", "x0", "xd")],
                       pb.max_vel, pb.max_acc)
    h_out = capi.HostResult(B, N)
    for name in ("rtn", "infeas_out", "line_failed_out", "iters", "cost", "x_final", "poly_coeff", "bez_coeff", "poly_time", "jerk", "stats"):
        t = pinned_like(getattr(h_out, name))
        keep["o_" + name] = t
        setattr(h_out, name, t.numpy())
    s0 = dict(w_snap=opts["w_snap0"], w_terminal=opts["w_terminal0"], w_time=opts["w_time0"], iter_max=opts["iter_max0"])
    s1 = dict(w_snap=opts["w_snap"], w_terminal=opts["w_terminal"], w_time=opts["w_time"], iter_max=opts["iter_max"])
    e2e_steps = max(3, min(args.steps, 5))
    solver.solve_two_stage(pbp, s0, s1, opts["time_power"], want_stage0=False, out1=h_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        flush.fill_(3)
        solver.solve_two_stage(pbp, s0, s1, opts["time_power"], want_stage0=False, out1=h_out)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    st_e2e = solver.stats()
    if world > 1:
        e2e_s = D.max_over_ranks(e2e_s, dev)
    e2e_value = world * B * e2e_steps / e2e_s

    # ---- accounting ---------------------------------------------------------------------------------------------
    stats = d_out["stats"].cpu().numpy()
    rtn = d_out["rtn"].cpu().numpy()
    iters1 = d_out["iters"].cpu().numpy()
    bwd_knots = int(st.bwd_knots)
    fwd_knots = int(st.fwd_knots)
    mean_planes = float(pb.nplanes.mean())
    flops = bwd_flops_per_knot(mean_planes) * bwd_knots
    achieved_tf = flops / (kernel_ms * 1e-3) / 1e12
    peak_tf = solver.fma_peak_tflops(args.precision)
    elem = 8 if args.precision == "fp64" else 4
    m_c = 6 * mean_planes + 55
    # algorithmic bytes per knot visit: backward reads x,u,s(,y) and writes the gains; a rollout reads x,u,s(,y),gains
    # and writes x,u,s(,y)   (SURVEY.md 8(d) "algorithmic bytes"; y only in stage 0)
    alg_bytes = elem * (bwd_knots * (19 + 1.5 * m_c + 100) + fwd_knots * (2 * 19 + 3.0 * m_c + 100))
    traffic, traffic_src = None, None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of the solve kernel, one ncu --set full capture (profiles/)
        prof = json.load(open(os.path.join(ROOT, "profiles", "latest_ncu.json")))
        if prof.get("workload") == f"{B}x{N} {args.kind} {args.precision}":
            traffic, traffic_src = prof["dram_bytes_per_launch"], prof["source"]
    except Exception:
        pass
    roofline = {"bound": "fp64_fma" if args.precision == "fp64" else "fp32_fma", "achieved": achieved_tf, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": "register-resident FMA microbenchmark run on this device by bench.py "
                               "(MEASURED_PEAKS.json has no CUDA-core FMA entry)",
                "flops_per_bwd_knot": bwd_flops_per_knot(mean_planes), "bwd_knots_per_launch": bwd_knots,
                "fwd_knots_per_launch": fwd_knots, "kernel_ms": kernel_ms,
                "hbm_algorithmic_GBps": alg_bytes / (kernel_ms * 1e-3) / 1e9}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        roofline["hbm_peak_GBps_measured"] = peaks.get("hbm_gbs")
    except Exception:
        pass
    cyc = stats[:, 4:7].sum(0)
    line = {
        "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64" if args.precision == "fp64" else "f32", "data": "synthetic", "config": workload_config(args, B),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": int(st_e2e.h2d_bytes),
                "d2h_bytes_per_step": int(st_e2e.d2h_bytes), "ms_per_step": e2e_s / e2e_steps * 1e3,
                "h2d_ms": st_e2e.h2d_ms, "d2h_ms": st_e2e.d2h_ms, "kernel_ms": st_e2e.kernel_ms},
        "gpu_launches": int(st.kernel_launches) * args.steps,
        "roofline": roofline,
        "solve_stats": {"converged_frac": float(np.isin(rtn, (1, 2)).mean()), "rtn_hist": {int(k): int(v) for k, v in zip(*np.unique(rtn, return_counts=True))},
                        "mean_stage1_iters": float(iters1.mean()), "bwd_sweeps_per_solve": st.bwd_sweeps / B,
                        "rollouts_per_solve": st.fwd_trials / B, "grid_blocks": st.grid_blocks,
                        "block_threads": st.block_threads, "smem_bytes_per_block": st.smem_bytes_per_block,
                        "warp_slots": st.workspace_slots, "coop_jobs": int(st.coop_jobs), "helper_units": int(st.helper_units),
                        "spec_searches": int(st.spec_searches), "spec_trials": int(st.spec_trials),
                        "cycle_share_bwd": float(cyc[0] / max(1, cyc[2])),
                        "cycle_share_linesearch": float(cyc[1] / max(1, cyc[2]))},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:   # reported at N = 1 only
        from oracle import oracle_py as O   # checker / CPU baseline only, never on the product path
        O.build(ref=False)
        cores = host_cores()
        smp = min(args.cpu_sample, B)
        sub = pb.slice(0, smp)
        t0 = time.perf_counter()
        _, a1 = O.two_stage_batch(sub, nthreads=cores)
        cpu_s = time.perf_counter() - t0
        g_cost = d_out["cost"][:smp].cpu().numpy()
        g_rtn = rtn[:smp]
        tol = 1e-5 if args.precision == "fp64" else 1e-3
        ok = (g_rtn == a1.rtn) & (np.abs(g_cost - a1.cost) <= tol * np.abs(a1.cost))
        line["cpu_baseline"] = {"value": smp / cpu_s, "unit": "solves/s", "cores": cores, "kind": "port",
                                "sample": f"first {smp} trajectories of rank 0's batch, two-stage protocol, {cores} OpenMP threads, "
                                          "fp64 C restatement of ddp_optimizer.cpp (oracle/ipddp_oracle.c)"}
        line["parity_sample"] = {"n": int(smp), "tol": tol, "frac_within_tol": float(ok.mean())}
        # Conditioning screen (tests/conftest.py): the trajectories on which the ORACLE moves by more than a tenth of the tolerance
        # under four 2^-48 relative input perturbations.  Parity is claimed on the rest; the screened ones are bounded by the
        # oracle's own spread (same return code as one of its five runs, cost inside their min..max widened by 10 %).
        nscr = min(args.screen_sample, smp)
        if nscr > 0:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from conftest import perturbed_batches
            ssub = pb.slice(0, nscr)
            runs = [a1] + [O.two_stage_batch(q, nthreads=cores)[1] for q in perturbed_batches(ssub)]
            cond = np.ones(nscr, dtype=bool)
            for r in runs[1:]:
                cond &= (r.rtn == a1.rtn[:nscr]) & (r.iters == a1.iters[:nscr]) & \
                        (np.abs(r.cost - a1.cost[:nscr]) <= 0.1 * tol * np.abs(a1.cost[:nscr])) & \
                        (np.max(np.abs(r.poly_time - a1.poly_time[:nscr]), axis=1) <= 0.1 * tol * np.max(np.abs(a1.poly_time[:nscr]), axis=1))
            cs = np.stack([r.cost[:nscr] for r in runs]); rs = np.stack([r.rtn[:nscr] for r in runs])
            lo, hi = cs.min(0), cs.max(0)
            inside = (g_cost[:nscr] >= lo - 0.1 * np.abs(lo)) & (g_cost[:nscr] <= hi + 0.1 * np.abs(hi)) & (rs == g_rtn[:nscr][None]).any(0)
            line["parity_sample"].update({
                "screen_n": int(nscr), "frac_screened": float((~cond).mean()),
                "frac_conditioned_within_tol": float(ok[:nscr][cond].mean()) if cond.any() else None,
                "screened_out_inside_oracle_spread": float(inside[~cond].mean()) if (~cond).any() else None,
                "screen": "oracle alone: 4 input perturbations of 2^-48 relative; screened = oracle cost / segment times move by > tol/10 or "
                          "its return code / iteration count changes"})
    if rank == 0 and world == 1 and not args.no_model_b:
        try:
            line["model_b_quadrotor12_fp32"] = model_b_report(capi, torch, dev, local_rank, flush, max(3, min(args.steps, 10)),
                                                              not args.no_cpu_baseline)
        except Exception as e:   # the headline line must not depend on the secondary report
            line["model_b_quadrotor12_fp32"] = {"error": repr(e)}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
