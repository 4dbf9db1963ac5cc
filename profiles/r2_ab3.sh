#!/bin/bash
# short A/B (one round) of the tuning builds build/tune/lib_*.so: loaded throughput (ab_one.py) and single-warp latency (lone_probe.py)
mkdir -p gpurun_out
: > gpurun_out/r2_ab3.txt
for f in build/tune/lib_*.so; do
  n=$(basename $f .so)
  SE3MPC_LIB=$PWD/$f timeout 300 python profiles/ab_one.py "$n" 2>&1 | tail -2 | tee -a gpurun_out/r2_ab3.txt
  echo "$n lone: $(SE3MPC_LIB=$PWD/$f timeout 120 python profiles/lone_probe.py 2>&1 | tail -1 | cut -c1-200)" | tee -a gpurun_out/r2_ab3.txt
done
