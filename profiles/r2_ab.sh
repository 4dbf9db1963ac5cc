#!/bin/bash
# A/B of the tuning builds build/tune/lib_*.so (optional per-library environment in build/tune/lib_X.env); two interleaved rounds
mkdir -p gpurun_out
for i in 1 2; do
for f in build/tune/lib_*.so; do
  envf="${f%.so}.env"; e=""; [ -f "$envf" ] && e=$(cat "$envf")
  env $e SE3MPC_LIB=$PWD/$f timeout 300 python profiles/ab_one.py "$(basename $f .so) $e" 2>&1 | tail -2 | tee -a gpurun_out/r2_ab.txt
done
done
