#!/bin/bash
# resident CTAs per SM (x4 warps) A/B through the SE3MPC_CTAS_PER_SM override: five bench batches at 4096, then 8192 and 16384
for c in 2 3; do
  echo "== SE3MPC_CTAS_PER_SM=$c"
  SE3MPC_CTAS_PER_SM=$c timeout 200 python profiles/lpt_bound.py | tail -1
  SE3MPC_CTAS_PER_SM=$c timeout 120 python profiles/ncu_target.py 8192 2
  SE3MPC_CTAS_PER_SM=$c timeout 120 python profiles/ncu_target.py 16384 2
done
