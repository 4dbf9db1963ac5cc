#!/bin/bash
# A/B timing of tuning builds (build/tune/lib_*.so): the five bench.py batches at 4096 problems (profiles/lpt_bound.py, first
# column) and one 16384-problem launch (profiles/ncu_target.py); two interleaved rounds, each run wrapped in a timeout
for i in 1 2; do
for f in build/tune/lib_*.so; do
  echo "== $f"
  SE3MPC_LIB=$PWD/$f timeout 200 python profiles/lpt_bound.py | tail -1
  SE3MPC_LIB=$PWD/$f timeout 120 python profiles/ncu_target.py 16384 2
done
done
