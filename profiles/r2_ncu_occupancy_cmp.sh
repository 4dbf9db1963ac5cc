set -x
export SE3MPC_LIB=$PWD/build/tune/lib_b_c4.so
SE3MPC_CTAS_PER_SM=4 python profiles/ncu_target.py 16384 1 > gpurun_out/plain_c4.log 2>&1 && \
SE3MPC_CTAS_PER_SM=4 ncu --set full --clock-control none --import-source on -k regex:se3mpc_solve_kernel -s 1 -c 1 -o gpurun_out/r2_c4_16384 python profiles/ncu_target.py 16384 1 > gpurun_out/ncu_c4.log 2>&1
SE3MPC_CTAS_PER_SM=2 python profiles/ncu_target.py 16384 1 > gpurun_out/plain_c2.log 2>&1 && \
SE3MPC_CTAS_PER_SM=2 ncu --set full --clock-control none --import-source on -k regex:se3mpc_solve_kernel -s 1 -c 1 -o gpurun_out/r2_c2_16384 python profiles/ncu_target.py 16384 1 > gpurun_out/ncu_c2.log 2>&1
tail -3 gpurun_out/plain_c4.log gpurun_out/plain_c2.log gpurun_out/ncu_c4.log gpurun_out/ncu_c2.log
ls -la gpurun_out/
