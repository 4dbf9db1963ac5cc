#!/bin/bash
# instruction-supply metrics of the solver kernel for each tuning build (B = 8192, one launch)
M=gpu__time_duration.sum,sm__cycles_active.avg,smsp__inst_executed.sum,sm__icc_requests.sum,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,gcc__average_cache_request_hit_rate.pct,l1tex__throughput.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed
for f in build/tune/lib_*.so; do
  echo "== $f"
  SE3MPC_LIB=$PWD/$f timeout 300 ncu --metrics $M --clock-control none -k regex:se3mpc_solve -s 1 -c 1 python profiles/ncu_target.py 8192 1 2>&1 | grep -E "^    [a-z]|^B "
done
