set -x
mkdir -p gpurun_out
python profiles/ncu_target.py 4096 1 > gpurun_out/plain_solve4096.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:se3mpc_solve_kernel -s 1 -c 1 -o gpurun_out/r2b_solve_4096 \
    python profiles/ncu_target.py 4096 1 > gpurun_out/ncu_solve4096.log 2>&1
tail -2 gpurun_out/plain_solve4096.log
