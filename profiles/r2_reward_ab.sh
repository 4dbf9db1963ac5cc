#!/bin/bash
for i in 1 2; do for f in build/tune/lib_*.so; do echo -n "$(basename $f .so): "; SE3MPC_LIB=$PWD/$f timeout 200 python profiles/reward_bench.py 2>&1 | tail -1; done; done | tee gpurun_out/r2_reward_ab.txt
