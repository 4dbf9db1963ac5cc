#!/bin/bash
# A/B timing of tuning builds (build/tune/lib_*.so) with profiles/ncu_target.py; each run is wrapped in a timeout
for f in build/tune/lib_*.so; do
  echo "== $f"
  SE3MPC_LIB=$PWD/$f timeout 120 python profiles/ncu_target.py 4096 3
  SE3MPC_LIB=$PWD/$f timeout 120 python profiles/ncu_target.py 16384 2
done
