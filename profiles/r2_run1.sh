set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2c_gputests.txt
bash profiles/r2_ab.sh > /dev/null 2>&1
python profiles/ncu_target.py 4096 1 > gpurun_out/plain_fac4096.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:se3mpc_solve_kernel -s 1 -c 1 -o gpurun_out/r2_fac_4096 python profiles/ncu_target.py 4096 1 > gpurun_out/ncu_fac4096.log 2>&1
cat gpurun_out/r2c_gputests.txt gpurun_out/r2_ab.txt gpurun_out/plain_fac4096.log
