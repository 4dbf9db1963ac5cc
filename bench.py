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
# DRAM bytes per problem of se3mpc_solve_kernel from the committed ncu capture (profiles/README.md): slot traffic (records, factors)
NCU_DRAM_BYTES_PER_PROBLEM = 26.8e6
NCU_REWARD_DRAM_BYTES_PER_TRAJECTORY = 6495.0     # se3mpc_reward_kernel, same source (5456 algorithmic)


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
    sc = sample_scenarios(batch, seed)
    p = O.default_params()
    t0 = time.perf_counter()
    r = O.objective_batch(p, sc["x0"], sc["goal"], sc["gate"], sc["tra_pos"], sc["tra_ang"], sc["t"], nthreads=nthreads)
    dt = time.perf_counter() - t0
    conv = int(np.sum(r["status"] <= 1))
    return conv, dt, O.num_threads()


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
                             "sample": f"{sample} problems of the configs[1] workload per step x {args.steps} steps"},
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import learningagileflight_se3_b200 as se3
    from learningagileflight_se3_b200 import ops

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
    W, K, B = max(args.warmup, 3), args.steps, args.batch
    p = se3.default_params()

    # per-rank shard of the synthetic batch (weak scaling: B problems per GPU), resident in HBM
    def dev_inputs(seed):
        sc = sample_scenarios(B, seed)
        return sc, [torch.as_tensor(sc[k], dtype=torch.float64, device=dev) for k in ("x0", "goal", "gate", "tra_pos", "tra_ang", "t")]

    steps_in = [dev_inputs(1000 * (rank + 1) + s) for s in range(W + K)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    side = torch.cuda.Stream(dev) if world > 1 else None
    keep = []

    def one_step(inp, gather=True):
        r = ops.objective(*inp, params=p)
        if world > 1 and gather:
            # the only exchange of the path: rewards + status to rank 0.  The gather runs on a side stream behind
            # an event, so the next batch's solves overlap it (the problems are independent): a straggler batch on one rank then
            # delays only that rank, not every rank at every step.  The timed region ends after the last gather has completed.
            payload = torch.stack([r["reward"], r["status"].to(torch.float64)], 1)
            ready = torch.cuda.Event()
            ready.record(stream)
            with torch.cuda.stream(side):
                side.wait_event(ready)
                out = [torch.empty_like(payload) for _ in range(world)] if rank == 0 else None
                dist.gather(payload, out, dst=0)
            keep.append((payload, out))
        return r

    try:
        dev_uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        dev_uuid = None
    for s in range(W):
        one_step(steps_in[s][1])
    barrier()
    clocks = ClockSampler(local, dev_uuid)
    clocks.start()
    time.sleep(0.3)                      # (sampler start-up stays out of the timed region)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches0 = se3.launch_count()
    results = []
    barrier()
    # Each step is timed on the device by its own CUDA-event pair (host-side launch jitter between steps -- Python, the caching
    # allocator, NVML -- is not part of the path and was measured at 5..750 ms per run on the 8-GPU boxes); the gathers of a
    # multi-GPU run overlap the following steps on a side stream, and what is still outstanding after the last step is timed too.
    for s in range(K):
        flush.zero_()                                                   # L2 flush between timed iterations (outside the event pair)
        ev[s][0].record(stream)
        results.append(one_step(steps_in[W + s][1]))
        ev[s][1].record(stream)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(stream)
    if side is not None:
        stream.wait_stream(side)                                        # the job ends when the last gather is done
    g1.record(stream)
    barrier()
    launches = se3.launch_count() - launches0
    clk = clocks.stop()
    t_ms = sum(a.elapsed_time(b) for a, b in ev) + g0.elapsed_time(g1)  # EXACTLY K steps, device time, max over ranks below
    print(f"[rank {rank}] per-step ms: {[round(a.elapsed_time(b), 2) for a, b in ev]}  + outstanding gather {g0.elapsed_time(g1):.2f} ms", file=sys.stderr)
    conv = sum(int((r["status"] <= 1).sum().item()) for r in results)
    flops = sum(algorithmic_flops(r["iters"].cpu().numpy().astype(np.float64), r["nfact"].cpu().numpy().astype(np.float64),
                                  r["nls"].cpu().numpy().astype(np.float64)) for r in results)
    stat = torch.tensor([t_ms, float(conv), flops, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = stat[0:1].clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        sums = stat[1:].clone(); dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        t_ms, conv, flops, launches = tmax.item(), sums[0].item(), sums[1].item(), int(sums[2].item())
    value = conv / (t_ms * 1e-3)

    # ---- end to end through the host-buffer C ABI (pinned host numpy in, numpy out) ----
    host_in = []
    for s in range(W + K):
        sc = steps_in[s][0]
        pinned = {k: torch.as_tensor(sc[k], dtype=torch.float64).pin_memory() for k in ("x0", "goal", "gate", "tra_pos", "tra_ang", "t")}
        host_in.append({k: v.numpy() for k, v in pinned.items()} | {"_keep": pinned})
    for s in range(2):
        ops.objective_host(*(host_in[s][k] for k in ("x0", "goal", "gate", "tra_pos", "tra_ang", "t")), params=p)
    barrier()
    t0 = time.perf_counter()
    conv_e2e = 0
    for s in range(K):
        r = ops.objective_host(*(host_in[W + s][k] for k in ("x0", "goal", "gate", "tra_pos", "tra_ang", "t")), params=p)
        conv_e2e += int(np.sum(r["status"] <= 1))
    barrier()
    t_e2e = time.perf_counter() - t0
    e2e_stat = torch.tensor([t_e2e, float(conv_e2e)], dtype=torch.float64, device=dev)
    if world > 1:
        tm = e2e_stat[0:1].clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        cs = e2e_stat[1:].clone(); dist.all_reduce(cs, op=dist.ReduceOp.SUM)
        t_e2e, conv_e2e = tm.item(), cs.item()
    h2d = B * (13 + 3 + 12 + 3 + 3 + 1) * 8
    d2h = B * (8 + 8 + 8 + 4 * 4)          # reward, cost, kkt_error + status/iters/nfact/nls

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
            a.record(stream)
            L.se3mpc_fp64_probe_device(1 << 13, C.c_void_p(sink.data_ptr()), C.c_void_p(stream.cuda_stream), C.byref(fl))
            b.record(stream)
            torch.cuda.synchronize(dev)
            best = max(best, fl.value / (a.elapsed_time(b) * 1e-3) / 1e12)
        achieved = flops / world / (t_ms * 1e-3) / 1e12 if world > 1 else flops / (t_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "se3mpc_solve_kernel", "achieved": achieved, "peak": best, "unit": "TFLOP/s",
                    "frac": achieved / best if best else None, "traffic": NCU_DRAM_BYTES_PER_PROBLEM * B,
                    "pipe": "FP64 tensor (DMMA.8x8x4) + FP64 FMA: same 16 FMA/clk/SMSP datapath",
                    "peak_source": "measured live: FP64 DMMA probe kernel (MEASURED_PEAKS.json has only the bf16 tensor figure; this path computes in f64)",
                    "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one 4096-problem launch (profiles/), scaled per problem",
                    "flops_per_step": flops / K / world, "model": "SURVEY.md 8(d): n_fact*W_fact + n_iter*(W_solve+W_eval) + n_ls*W_ls per problem"}
        # stand-alone reward kernel (HBM bound): 5304+120 B read + 32 B written per trajectory
        NB = 1 << 18
        Xr = torch.randn(NB, 51, 13, dtype=torch.float64, device=dev)
        gate_r = steps_in[0][1][2][:1].expand(NB, 12).contiguous()
        goal_r = steps_in[0][1][1][:1].expand(NB, 3).contiguous()
        for _ in range(3):
            ops.reward(Xr, gate_r, goal_r, params=p)
        tr = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); ops.reward(Xr, gate_r, goal_r, params=p); b.record(stream)
            torch.cuda.synchronize(dev)
            tr.append(a.elapsed_time(b))
        rb = NB * (51 * 13 * 8 + 15 * 8 + 32)
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        ach_r = rb / (min(tr) * 1e-3) / 1e9
        roofline_reward = {"bound": "hbm", "kernel": "se3mpc_reward_kernel", "achieved": ach_r, "peak": hbm_peak, "unit": "GB/s",
                           "frac": ach_r / hbm_peak, "traffic": NCU_REWARD_DRAM_BYTES_PER_TRAJECTORY * NB, "trajectories": NB,
                           "traffic_source": "ncu --set full of one 2^18-trajectory launch (profiles/r1_ncu_full_raw_reward_kernel.csv): 16-byte-granular bulk copies of 8-byte-aligned 416 B chunks over-fetch ~18%",
                           "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s"}
        del Xr
        cpu_b = None
        if not args.no_cpu_baseline:
            conv_c, dt_c, cores = cpu_reference(args.cpu_sample, 1)
            cpu_b = {"value": conv_c / dt_c, "unit": UNIT, "cores": cores, "kind": "port",
                     "sample": f"{args.cpu_sample} problems of the same workload (seed 1), {dt_c:.1f} s wall on {cores} threads; "
                               "restated fp64 CPU IPM + reward (oracle/), not CasADi/IPOPT"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": t_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "batch_per_gpu": B, "horizon": 50, "parallelism": f"problems sharded x{world}",
                           "l2_flush": "256 MiB zero-fill between timed steps (outside the CUDA-event pairs)",
                           "converged_fraction": conv / (K * B * world)},
                "clocks": clk, "gpu_launches": int(launches),
                "e2e": {"value": conv_e2e / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "api": "ops.objective_host -> se3mpc_objective_host (pinned host numpy in, host numpy out)"},
                "roofline": roofline, "roofline_reward": roofline_reward, "cpu_baseline": cpu_b}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
