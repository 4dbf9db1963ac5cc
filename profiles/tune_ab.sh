#!/bin/bash
# A/B timing of tuning builds (build/tune/lib_<warps per CTA>_<CTAs per SM>.so) with profiles/ncu_target.py
for f in build/tune/lib_*.so; do
  echo "== $f"
  SE3MPC_LIB=$PWD/$f python profiles/ncu_target.py 4096 2
  SE3MPC_LIB=$PWD/$f python profiles/ncu_target.py 16384 1
done
