"""One A/B arm: the five bench.py batches (seeds 1003..1007, 4096 problems, best of 2) and one 16384-problem launch of the library
named by SE3MPC_LIB (tuning builds under build/tune/).  Prints one summary line."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learningagileflight_se3_b200 as se3  # noqa: E402
from tests.common import sample_scenarios  # noqa: E402

label = sys.argv[1] if len(sys.argv) > 1 else "arm"
big = int(sys.argv[2]) if len(sys.argv) > 2 else 16384


def timed(args, reps):
    best, r = 1e9, None
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = se3.ops.objective(*args); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, r


def dev(sc):
    return [torch.as_tensor(sc[k], dtype=torch.float64, device="cuda") for k in ("x0", "goal", "gate", "tra_pos", "tra_ang", "t")]


warm = dev(sample_scenarios(2048, 5))
timed(warm, 1)
ms5, conv, chk = [], 0, 0.0
for seed in range(1003, 1008):
    ms, r = timed(dev(sample_scenarios(4096, seed)), 2)
    ms5.append(round(ms, 1)); conv += int((r["status"] <= 1).sum()); chk += float(r["reward"].sum())
msb, rb = timed(dev(sample_scenarios(big, 1)), 3) if big > 0 else (0.0, None)      # best of 3: the first big launch of a process is bimodal
print(f"{label:28s} 4096x5 ms {ms5} sum {sum(ms5):.1f} -> {5 * 4096 / sum(ms5):.1f} k/s conv {conv} rewsum {chk:.6f} | "
      f"{big}: {msb:.1f} ms -> {big / max(msb, 1e-9):.1f} k/s", flush=True)
# the same five batches pipelined over two streams (each stream owns its workspace): the tail of one batch overlaps the head of the next
ins = [dev(sample_scenarios(4096, seed)) for seed in range(1003, 1008)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
for rep in range(2):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for st in streams:
        st.wait_stream(torch.cuda.current_stream())
    outs = []
    for i, x in enumerate(ins):
        with torch.cuda.stream(streams[i & 1]):
            outs.append(se3.ops.objective(*x))
    for st in streams:
        torch.cuda.current_stream().wait_stream(st)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
print(f"{label:28s} 4096x5 on 2 streams: {ms:.1f} ms -> {5 * 4096 / ms:.1f} k/s", flush=True)
