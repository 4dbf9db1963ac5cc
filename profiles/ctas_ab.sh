#!/bin/bash
for v in auto 2 3; do
  echo "== SE3MPC_CTAS_PER_SM=$v"
  if [ $v = auto ]; then unset SE3MPC_CTAS_PER_SM; else export SE3MPC_CTAS_PER_SM=$v; fi
  python profiles/ncu_target.py 4096 3; python profiles/ncu_target.py 4096 3; python profiles/ncu_target.py 16384 2
done
