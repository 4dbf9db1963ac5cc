set -x
export SE3MPC_LIB=$PWD/build/tune/lib_c_l1diet.so
python profiles/ncu_target.py 148 1 > gpurun_out/plain_lone.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:se3mpc_solve_kernel -s 1 -c 1 -o gpurun_out/r2_lone_148 python profiles/ncu_target.py 148 1 > gpurun_out/ncu_lone.log 2>&1
python profiles/ncu_target.py 16384 1 > gpurun_out/plain_diet.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:se3mpc_solve_kernel -s 1 -c 1 -o gpurun_out/r2_diet_16384 python profiles/ncu_target.py 16384 1 > gpurun_out/ncu_diet.log 2>&1
cat gpurun_out/plain_lone.log gpurun_out/plain_diet.log | tail -4
