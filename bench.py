#!/usr/bin/env python
"""bench.py -- converged SE(3) MPC solves/s (BASELINE.json metric).

A "step" is one pass of the hot path over one batch: BASELINE.json configs[1], "batch of 4096 random initial
conditions: forward MPC solve + solid_geometry traversal reward" = run_quad.objective x 4096 per GPU, as ONE
se3mpc_objective launch.  Inputs are synthetic (the reference's nn_sample rules, seeded).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference      # the reference arm: CPU path on the host cores

Prints ONE JSON line.  `value` = converged solves/s with inputs resident in HBM (CUDA events on the launching
stream, max over ranks); `e2e` = the same through the host-buffer C ABI (pinned host numpy in, host numpy out);
`roofline` = the solver kernel against the FP64 tensor/FMA peak (measured live by a DMMA probe kernel since
MEASURED_PEAKS.json carries no fp64 figure), `roofline_reward` = the stand-alone reward kernel against measured
HBM bandwidth; `cpu_baseline` = the CPU oracle ("port": restated fp64 IPM, NOT CasADi/IPOPT, which is not
installable offline) on all host cores over a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tests.common import sample_scenarios  # noqa: E402

METRIC = "converged SE(3) MPC solves/sec"
UNIT = "solves/s"
WORKLOAD = "configs[1]: 4096 random initial conditions per GPU, run_quad.objective = MPC solve (N=50, nx=13, nu=4, tol 1e-8, cold start) + gate-traversal reward"

# algorithmic flops per problem from the kernel's own counters (textbook dense-block Riccati at nx=13, nu=4)
N_H = 50
W_FACT = N_H * 8267.0
W_SOLVE = N_H * 1124.0
W_EVAL = N_H * 600.0
W_LS = N_H * 150.0


def ncu_dram_bytes_per_unit(pattern):
    """DRAM traffic per problem (or trajectory) of a kernel from the newest committed `ncu --set full` raw page matching
    profiles/<pattern> (the file name ends in _B<units>.csv): dram__bytes_read.sum + dram__bytes_write.sum of that launch divided by
    the units it processed.  Returns (bytes_per_unit, file name) or (None, None): the figure is read from the capture, not hard-coded."""
    import csv
    import glob
    import re
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", pattern))):
        m = re.search(r"_B(\d+)\.csv$", f)
        if not m:
            continue
        try:
            rows = list(csv.reader(open(f)))
            hdr, units, vals = rows[0], rows[1], rows[2]
            tot = 0.0
            for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                i = hdr.index(name)
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[units[i]]
                tot += float(vals[i]) * scale
            best = (tot / int(m.group(1)), os.path.basename(f))
        except Exception:
            continue
    return best if best else (None, None)


def algorithmic_flops(iters, nfact, nls):
    return float(np.sum(nfact * W_FACT + iters * (W_SOLVE + W_EVAL) + nls * W_LS))


class ClockSampler:
    """SM clock / throttle reasons during the timed region (the clocks line of B200_PROFILING.md), sampled in-process through
    NVML (nvidia_ml_py): a spawned `nvidia-smi -lms` blocks the CUDA launch path of the benchmark for tens to hundreds of ms
    per sample on an 8-GPU box.  Falls back to nvidia-smi when the NVML binding is missing."""

    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index, uuid=None):
        self.index, self.uuid, self.rows, self.proc, self.stop_flag, self.thread, self.nv = index, uuid, [], None, False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(self.uuid)).encode() if not str(self.uuid).startswith("GPU-") else str(self.uuid).encode())
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nv = (pynvml, h)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nv = None
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", os.environ.get("SE3_SMI_MS", "100")],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        pynvml, h = self.nv
        period = float(os.environ.get("SE3_SMI_MS", "50")) * 1e-3
        while not self.stop_flag:
            try:
                mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    mask = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                flags = ["Active" if mask & bit else "Not Active" for bit, _ in self.REASONS]
                self.rows.append([str(mhz), str(self.max_mhz), "0"] + flags)
            except Exception:
                pass
            time.sleep(period)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [v for v in sm if v > 0.5 * mx] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvml" if self.nv else "nvidia-smi"}


def cpu_reference(batch, seed, nthreads=0):
    """The reference arm's engine: the CPU oracle (restated fp64 IPM + reward), one problem per host thread."""
    from oracle import oracle as O
    O.build()
    global CPU_FLAGS
    CPU_FLAGS = O.use_native()          # compiled for the host cores it is timed on
    sc = sample_scenarios(batch, seed)
    p = O.default_params()
    t0 = time.perf_counter()
    r = O.objective_batch(p, sc["x0"], sc["goal"], sc["gate"], sc["tra_pos"], sc["tra_ang"], sc["t"], nthreads=nthreads)
    dt = time.perf_counter() - t0
    conv = int(np.sum(r["status"] <= 1))
    return conv, dt, O.num_threads()


CPU_FLAGS = ""


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.ref_batch
    for _ in range(args.warmup):
        cpu_reference(min(sample, 32), 999)
    conv_total, t_total, cores = 0, 0.0, 1
    for s in range(args.steps):
        conv, dt, cores = cpu_reference(sample, 1 + s)
        conv_total += conv; t_total += dt
    val = conv_total / t_total
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample_per_step": sample,
                       "note": "CasADi/IPOPT (the reference's solver) is not installable offline; this arm times the repo's CPU restatement "
                               "of the same NLP + IPOPT-style interior point + reward on all host cores"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} problems of the configs[1] workload per step x {args.steps} steps", "build": "gcc " + CPU_FLAGS},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4096, help="problems per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-batch", type=int, default=512)
    ap.add_argument("--cpu-sample", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=2, help="steps in flight: step s runs on stream s %% streams (1 = serial)")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra.configs measurements (latency, RL step, closed loop, 1M sweep)")
    ap.add_argument("--sweep", type=int, default=1_000_000, help="problems of the configs[4] sweep (total over all GPUs)")
    ap.add_argument("--episodes", type=int, default=4096, help="episodes of the configs[3] closed loop (total over all GPUs)")
    ap.add_argument("--ticks", type=int, default=50)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import hashlib
    from concurrent.futures import ThreadPoolExecutor

    import torch
    import torch.distributed as dist
    import learningagileflight_se3_b200 as se3
    from learningagileflight_se3_b200 import dist as sdist
    from learningagileflight_se3_b200 import learning, ops, quad_moving

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner straight to fd 1 on first use; keep stdout for the ONE JSON line.
        sys.stdout.flush()
        keep = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(keep, 1)
            os.close(keep)
    W, K, B, S = max(args.warmup, 3), args.steps, args.batch, max(1, args.streams)
    p = se3.default_params()
    KEYS = ("x0", "goal", "gate", "tra_pos", "tra_ang", "t")
    f64 = dict(dtype=torch.float64, device=dev)

    # per-rank shard of the synthetic batch (weak scaling: B problems per GPU), resident in HBM
    def dev_inputs(seed):
        sc = sample_scenarios(B, seed)
        return sc, [torch.as_tensor(sc[k], **f64) for k in KEYS]

    steps_in = [dev_inputs(1000 * (rank + 1) + s) for s in range(W + K)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    main = torch.cuda.current_stream(dev)
    lanes = [torch.cuda.Stream(dev) for _ in range(S)] if S > 1 else [main]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    side = torch.cuda.Stream(dev) if world > 1 else None
    keep = []

    def one_step(inp, lane):
        r = ops.objective(*inp, params=p)
        if world > 1:
            # the only exchange of the path: rewards + status to rank 0.  The gather runs on a side stream behind
            # an event, so the following batches' solves overlap it (the problems are independent): a straggler batch on one rank then
            # delays only that rank, not every rank at every step.  The timed region ends after the last gather has completed.
            payload = torch.stack([r["reward"], r["status"].to(torch.float64)], 1)
            ready = torch.cuda.Event()
            ready.record(lane)
            payload.record_stream(side)
            with torch.cuda.stream(side):
                side.wait_event(ready)
                out = [torch.empty_like(payload) for _ in range(world)] if rank == 0 else None
                dist.gather(payload, out, dst=0)
            keep.append((payload, out))
        return r

    def run_steps(first, count, timed):
        """`count` steps through the pipeline: step s on stream s % S (each stream owns its solver workspace), so that the straggler tail
        of one batch is filled by the next batch instead of idling the GPU.  Returns (results, per-step event pairs, total ms)."""
        ev, res = [], []
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(main)
        for ln in lanes:
            if ln is not main:
                ln.wait_stream(main)
        for s in range(count):
            ln = lanes[s % S]
            with torch.cuda.stream(ln):
                flush.zero_()                                           # L2 flush before every step (on the step's stream)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(ln)
                res.append(one_step(steps_in[first + s][1], ln))
                b.record(ln)
                ev.append((a, b))
        for ln in lanes:
            if ln is not main:
                main.wait_stream(ln)
        if side is not None:
            main.wait_stream(side)                                      # the job ends when the last gather is done
        t1.record(main)
        return res, ev, (t0, t1)

    try:
        dev_uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        dev_uuid = None
    run_steps(0, W, False)
    barrier()
    clocks = ClockSampler(local, dev_uuid)
    clocks.start()
    time.sleep(0.3)                      # (sampler start-up stays out of the timed region)
    launches0 = se3.launch_count()
    barrier()
    # EXACTLY K steps, bracketed by barrier + synchronize on both sides; device time by CUDA events on the launching stream (the lanes
    # fork from / join into it), max over ranks below.
    results, ev, (t0, t1) = run_steps(W, K, True)
    barrier()
    launches = se3.launch_count() - launches0
    clk = clocks.stop()
    t_ms = t0.elapsed_time(t1)
    step_ms = [round(a.elapsed_time(b), 2) for a, b in ev]
    print(f"[rank {rank}] {K} steps on {S} stream(s): {t_ms:.2f} ms; per-step (overlapping) ms: {step_ms}", file=sys.stderr)
    conv = sum(int((r["status"] <= 1).sum().item()) for r in results)
    hist = np.sum([np.bincount(r["status"].cpu().numpy(), minlength=5)[:5] for r in results], 0).astype(np.float64)
    flops = sum(algorithmic_flops(r["iters"].cpu().numpy().astype(np.float64), r["nfact"].cpu().numpy().astype(np.float64),
                                  r["nls"].cpu().numpy().astype(np.float64)) for r in results)
    # the same K batches one after the other on ONE stream: the time of the solver kernel itself (roofline) and the step-to-step spread
    serial_ms = []
    ops.objective(*steps_in[0][1], params=p); torch.cuda.synchronize(dev)       # (this stream's workspace is allocated outside the timing)
    for s in range(K):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(main); ops.objective(*steps_in[W + s][1], params=p); b.record(main)
        torch.cuda.synchronize(dev)
        serial_ms.append(a.elapsed_time(b))
    stat = torch.tensor([t_ms, float(conv), flops, float(launches), sum(serial_ms)] + hist.tolist(), dtype=torch.float64, device=dev)
    if world > 1:
        tmax = stat[[0, 4]].clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        sums = stat.clone(); dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        t_ms, t_serial = tmax[0].item(), tmax[1].item()
        conv, flops, launches = sums[1].item(), sums[2].item(), int(sums[3].item())
        hist = sums[5:].cpu().numpy()
    else:
        t_serial = sum(serial_ms)
    value = conv / (t_ms * 1e-3)

    # ---- end to end through the host-buffer C ABI (pinned host numpy in, numpy out), S host threads = S batches in flight ----
    host_in = []
    for s in range(W + K):
        sc = steps_in[s][0]
        pinned = {k: torch.as_tensor(sc[k], dtype=torch.float64).pin_memory() for k in KEYS}
        host_in.append({k: v.numpy() for k, v in pinned.items()} | {"_keep": pinned})

    def host_step(s):
        with torch.cuda.device(dev):
            return ops.objective_host(*(host_in[s][k] for k in KEYS), params=p)

    pool = ThreadPoolExecutor(max_workers=S)
    list(pool.map(host_step, range(min(W, 2 * S))))
    barrier()
    t0w = time.perf_counter()
    host_res = list(pool.map(host_step, range(W, W + K)))
    conv_e2e = sum(int(np.sum(r["status"] <= 1)) for r in host_res)
    gather_bytes = 0
    if world > 1:          # results to rank 0: one H2D of the K x B (reward, status) pairs + one NCCL gather, inside the timed region
        pay = torch.as_tensor(np.stack([np.concatenate([r["reward"] for r in host_res]),
                                        np.concatenate([r["status"] for r in host_res]).astype(np.float64)], 1)).to(dev)
        out = [torch.empty_like(pay) for _ in range(world)] if rank == 0 else None
        dist.gather(pay, out, dst=0)
        gather_bytes = pay.numel() * 8
    barrier()
    t_e2e = time.perf_counter() - t0w
    pool.shutdown()
    e2e_stat = torch.tensor([t_e2e, float(conv_e2e)], dtype=torch.float64, device=dev)
    if world > 1:
        tm = e2e_stat[0:1].clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        cs = e2e_stat[1:].clone(); dist.all_reduce(cs, op=dist.ReduceOp.SUM)
        t_e2e, conv_e2e = tm.item(), cs.item()
    h2d = B * (13 + 3 + 12 + 3 + 3 + 1) * 8 + (gather_bytes // K if world > 1 else 0)
    d2h = B * (8 + 8 + 8 + 4 * 4)          # reward, cost, kkt_error + status/iters/nfact/nls

    # ---- shard-count invariance: ONE fixed-seed global batch, sharded by problem index through dist.sharded_call, gathered to rank 0.
    # Every problem is solved by one warp with a fixed arithmetic order, so the hash must be identical for 1, 2, 4 and 8 GPUs.
    def objective_shard(**kw):
        r = ops.objective(*(kw[k] for k in KEYS), params=p)
        return {"reward": r["reward"], "status": r["status"], "iters": r["iters"]}

    def sha(res):
        h = hashlib.sha256()
        h.update(np.ascontiguousarray(res["reward"].cpu().numpy()).tobytes())
        h.update(np.ascontiguousarray(res["status"].cpu().numpy().astype(np.int32)).tobytes())
        return h.hexdigest()

    inv = sdist.sharded_call(objective_shard, {k: v for k, v in sample_scenarios(4096, 77).items() if k in KEYS}, device=dev, keys=("reward", "status"))
    invariance_sha = sha(inv) if rank == 0 else None

    extra = {}
    if not args.no_extras:
        extra = run_extras(args, dev, rank, world, p, KEYS, f64, se3, ops, sdist, learning, quad_moving, dist, barrier, objective_shard, sha)

    line = None
    if rank == 0:
        # ---- roofline denominators measured live ----
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        import ctypes as C
        sink = torch.zeros(8, dtype=torch.float64, device=dev)
        fl = C.c_double(0)
        L = se3._lib.lib()
        best = 0.0
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(main)
            L.se3mpc_fp64_probe_device(1 << 13, C.c_void_p(sink.data_ptr()), C.c_void_p(main.cuda_stream), C.byref(fl))
            b.record(main)
            torch.cuda.synchronize(dev)
            best = max(best, fl.value / (a.elapsed_time(b) * 1e-3) / 1e12)
        # achieved = algorithmic flops per launch / average launch duration of the solver kernel timed alone (the serial pass above)
        achieved = (flops / world / K) / (t_serial / K * 1e-3) / 1e12
        per_problem, src = ncu_dram_bytes_per_unit("r*_ncu_full_raw_solve_B*.csv")
        roofline = {"bound": "tensor", "kernel": "se3mpc_solve_kernel", "achieved": achieved, "peak": best, "unit": "TFLOP/s",
                    "frac": achieved / best if best else None, "traffic": per_problem * B if per_problem else None,
                    "pipe": "FP64 tensor (DMMA.8x8x4) + FP64 FMA: same 16 FMA/clk/SMSP datapath",
                    "peak_source": "measured live: FP64 DMMA probe kernel (MEASURED_PEAKS.json has only the bf16 tensor figure; this path computes in f64)",
                    "traffic_source": f"profiles/{src}: dram__bytes_read.sum + dram__bytes_write.sum of that ncu --set full capture, per problem x {B}" if src else None,
                    "launch_ms": t_serial / K, "flops_per_step": flops / K / world,
                    "frac_pipelined": (flops / world) / (t_ms * 1e-3) / 1e12 / best if best else None,
                    "model": "SURVEY.md 8(d): n_fact*W_fact + n_iter*(W_solve+W_eval) + n_ls*W_ls per problem"}
        # stand-alone reward kernel (HBM bound): 5304+120 B read + 32 B written per trajectory
        NB = 1 << 18
        Xr = torch.randn(NB, 51, 13, **f64)
        gate_r = steps_in[0][1][2][:1].expand(NB, 12).contiguous()
        goal_r = steps_in[0][1][1][:1].expand(NB, 3).contiguous()
        for _ in range(3):
            ops.reward(Xr, gate_r, goal_r, params=p)
        tr = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(main); ops.reward(Xr, gate_r, goal_r, params=p); b.record(main)
            torch.cuda.synchronize(dev)
            tr.append(a.elapsed_time(b))
        rb = NB * (51 * 13 * 8 + 15 * 8 + 32)
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        ach_r = rb / (min(tr) * 1e-3) / 1e9
        per_traj, src_r = ncu_dram_bytes_per_unit("r*_ncu_full_raw_reward_B*.csv")
        roofline_reward = {"bound": "hbm", "kernel": "se3mpc_reward_kernel", "achieved": ach_r, "peak": hbm_peak, "unit": "GB/s",
                           "frac": ach_r / hbm_peak, "traffic": per_traj * NB if per_traj else None, "trajectories": NB,
                           "traffic_source": f"profiles/{src_r}" if src_r else None,
                           "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s"}
        del Xr
        cpu_b = None
        if not args.no_cpu_baseline:
            conv_c, dt_c, cores = cpu_reference(args.cpu_sample, 1)
            cpu_b = {"value": conv_c / dt_c, "unit": UNIT, "cores": cores, "kind": "port",
                     "sample": f"{args.cpu_sample} problems of the same workload (seed 1), {dt_c:.1f} s wall on {cores} threads; "
                               "restated fp64 CPU IPM + reward (oracle/), not CasADi/IPOPT", "build": "gcc " + CPU_FLAGS}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": t_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "batch_per_gpu": B, "horizon": 50, "parallelism": f"problems sharded x{world}",
                           "steps_in_flight": S,
                           "overlap": f"step s is launched on stream s % {S}: independent batches overlap, the straggler tail of one launch is filled by the next" if S > 1 else "none (one stream)",
                           "l2_flush": "256 MiB zero-fill before every timed step, on the step's stream; the solver workspace (370 MB per stream) also exceeds L2",
                           "converged_fraction": conv / (K * B * world)},
                "clocks": clk, "gpu_launches": int(launches),
                "status_hist": {name: int(v) for name, v in zip(("converged", "acceptable", "max_iter", "failed", "nan"), hist)},
                "invariance_sha": invariance_sha,
                "serial": {"value": conv / (t_serial * 1e-3) if t_serial else None, "unit": UNIT, "per_step_ms": [round(v, 2) for v in serial_ms],
                           "max_over_min": max(serial_ms) / min(serial_ms), "note": "the same K batches one at a time on one stream (rank 0's step times)"},
                "e2e": {"value": conv_e2e / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "api": f"ops.objective_host -> se3mpc_objective_host (pinned host numpy in, host numpy out), {S} host threads = {S} batches in flight"
                               + ("; + H2D and NCCL gather of (reward, status) to rank 0" if world > 1 else "")},
                "roofline": roofline, "roofline_reward": roofline_reward, "cpu_baseline": cpu_b, "extra": {"configs": extra}}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_extras(args, dev, rank, world, p, KEYS, f64, se3, ops, sdist, learning, quad_moving, dist, barrier, objective_shard, sha):
    """The other BASELINE.json configs, measured in the same run (rank 0 reports).  Device time by CUDA events where one launch is
    timed, wall clock between synchronisations for the multi-launch loops."""
    import torch
    out = {}
    main = torch.cuda.current_stream(dev)

    def maxf(v):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def sumf(v):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.item()

    # ---- configs[0]: single gate-traversal solve (quad_policy planner, one problem per call), rank 0 ----
    if rank == 0:
        sc = sample_scenarios(32, 0)
        ms_dev, ms_host, its = [], [], []
        for i in range(32):
            one = [torch.as_tensor(sc[k][i:i + 1], **f64) for k in KEYS]
            if i == 0:
                ops.objective(*one, params=p); torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(main); r = ops.objective(*one, params=p); b.record(main)
            torch.cuda.synchronize(dev)
            ms_dev.append(a.elapsed_time(b)); its.append(int(r["iters"][0]))
            t0 = time.perf_counter()
            ops.objective_host(*(sc[k][i:i + 1] for k in KEYS), params=p)
            ms_host.append((time.perf_counter() - t0) * 1e3)
        out["single_solve_latency_ms"] = {"workload": "configs[0]: run_quad.objective for ONE problem (32 different problems, one launch each)",
                                          "mean": float(np.mean(ms_dev)), "p50": float(np.median(ms_dev)), "p99": float(np.max(ms_dev)),
                                          "host_api_mean": float(np.mean(ms_host)), "host_api_p99": float(np.max(ms_host)),
                                          "mean_iterations": float(np.mean(its)), "ms_per_iteration": float(np.sum(ms_dev) / np.sum(its))}

    # ---- configs[2]: RL step, 1024 samples x 9 perturbed problems (run_quad.sol_gradient), one launch; samples sharded over the ranks ----
    nS = 1024
    lo, hi = sdist.shard_range(nS, rank, world)
    sc = sample_scenarios(nS, 2)
    sg = [torch.as_tensor(sc[k][lo:hi], **f64) for k in KEYS]
    ops.sol_gradient(*sg, params=p); barrier()
    best, conv9 = 1e9, 0
    for _ in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(main); g, st9 = ops.sol_gradient(*sg, params=p); b.record(main)
        torch.cuda.synchronize(dev)
        best = min(best, a.elapsed_time(b)); conv9 = int((st9 <= 1).sum())
    ms, conv9 = maxf(best), sumf(conv9)
    out["rl_step"] = {"workload": "configs[2]: 1024 samples x 9 perturbed solves + rewards + clip/scale epilogue, one se3mpc_sol_gradient launch per GPU",
                      "problems": 9 * nS, "ms": ms, "solves_per_s": conv9 / (ms * 1e-3), "pseudo_gradient_samples_per_s": nS / (ms * 1e-3),
                      "converged_fraction": conv9 / (9 * nS)}

    # ---- configs[3]: moving narrow gate, closed loop with the reference's shipped DNN-2 (nn3_1.pth weights as a fixture) ----
    E, ticks = args.episodes, args.ticks
    lo, hi = sdist.shard_range(E, rank, world)
    wts = np.load(os.path.join(ROOT, "tests", "golden", "nn3_1_weights.npz"))
    model = learning.dnn_from_arrays(wts, dev)
    sc = sample_scenarios(E, 3)
    x0, fin, g0 = (torch.as_tensor(sc[k][lo:hi], **f64) for k in ("x0", "goal", "gate"))
    velo = torch.tensor([1.0, 0.3, 0.4], **f64).expand(hi - lo, 3).contiguous()          # main.py:45
    gen = torch.Generator(device=dev); gen.manual_seed(3)
    noise = torch.clamp(0.1 * torch.randn(ticks * 10, E, 3, generator=gen, **f64), -0.1, 0.1)[:, lo:hi].contiguous()      # gate.move, quad_model.py:778
    quad_moving.run_episodes(model, x0[:64], fin[:64], g0[:64], velo[:64], math.pi / 2, n_ticks=2, gate_noise=noise[:, :64].contiguous(), groups=2)
    cl = {}
    for name, warm in (("cold_start", False), ("warm_start", True)):
        barrier()
        l0 = se3.launch_count()
        t0 = time.perf_counter()
        res = quad_moving.run_episodes(model, x0, fin, g0, velo, math.pi / 2, n_ticks=ticks, gate_noise=noise, warm_start=warm, groups=4)
        torch.cuda.synchronize(dev)
        dt = maxf(time.perf_counter() - t0)
        conv = sumf(int((res["status"] <= 1).sum()))
        sth = torch.bincount(res["status"].reshape(-1).to(torch.int64), minlength=5)[:5].to(torch.float64)
        if world > 1:
            dist.all_reduce(sth, op=dist.ReduceOp.SUM)
        cl[name] = {"seconds": dt, "solves_per_s": conv / dt, "converged_fraction": conv / (E * ticks), "ms_per_tick": dt / ticks * 1e3,
                    "status_hist": [int(v) for v in sth.cpu().numpy()],
                    "mean_iterations": sumf(float(res["iters"].double().sum())) / (E * ticks), "kernel_launches_per_tick_per_gpu": (se3.launch_count() - l0) / ticks,
                    "traversal_time_unconverged": sumf(int(res["tt_unconverged"].sum()))}
    out["closed_loop"] = {"workload": f"configs[3]: {E} episodes x {ticks} MPC ticks (main.py:65-116), DNN-2 = the reference's nn3_1.pth (fp32, weights from "
                                      "tests/golden/nn3_1_weights.npz), gate v = [1, 0.3, 0.4] + clipped N(0, 0.1) noise, w = pi/2; episodes sharded over the GPUs, 4 stream "
                                      "groups per GPU; cold_start = the reference's behaviour (fresh run_quad per tick), warm_start = previous plan shifted by one stage",
                          **cl}

    # ---- configs[4]: 1M independent traversal solves, sharded by problem index over the ranks, results gathered to rank 0 ----
    n = args.sweep
    lo, hi = sdist.shard_range(n, rank, world)
    CH = 62500                                          # problems are generated in chunks from (seed, chunk): any rank can build any part
    parts = []
    for ci in range(lo // CH, (hi + CH - 1) // CH if hi > lo else lo // CH):
        scc = sample_scenarios(min(CH, n - ci * CH), 400 + ci)
        a, b = max(lo, ci * CH) - ci * CH, min(hi, (ci + 1) * CH) - ci * CH
        parts.append({k: scc[k][a:b] for k in KEYS})
    shard = {k: torch.as_tensor(np.concatenate([q[k] for q in parts]) if parts else np.zeros((0,) + sample_scenarios(1, 0)[k].shape[1:]), **f64) for k in KEYS}
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record(main)
    res = sdist.sharded_call(objective_shard, shard, device=dev, keys=("reward", "status", "iters"), n_total=n)
    b.record(main)
    torch.cuda.synchronize(dev)
    my_s = a.elapsed_time(b) * 1e-3
    barrier()
    wall = maxf(time.perf_counter() - t0)
    per_rank = torch.zeros(world, dtype=torch.float64, device=dev)
    per_rank[rank] = my_s
    if world > 1:
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
    if rank == 0:
        st = res["status"].cpu().numpy()
        conv = int((st <= 1).sum())
        secs = float(per_rank.max())
        out["sweep_1M"] = {"workload": f"configs[4]: {n} independent run_quad.objective problems, contiguous shards by problem index over {world} GPU(s) "
                                       "(dist.sharded_call: one launch per GPU, (reward, status, iters) gathered to rank 0)",
                           "problems": n, "seconds": secs, "wall_seconds": wall, "solves_per_s": conv / secs, "converged_fraction": conv / n,
                           "per_rank_seconds": [round(float(v), 3) for v in per_rank.cpu().numpy()],
                           "imbalance_max_over_mean": float(per_rank.max() / per_rank.mean()),
                           "status_hist": np.bincount(st, minlength=5)[:5].tolist(), "mean_iterations": float(res["iters"].double().mean()),
                           "sha256": sha(res)}
    return out


if __name__ == "__main__":
    main()
