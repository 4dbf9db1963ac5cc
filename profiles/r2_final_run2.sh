# round-2 artefacts of the final build (after the branch-free factor sweep): GPU tests, smoke, bench lines, launch list, full capture of the solver
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2_final_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/r2_final_gputests.txt 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench.err
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python profiles/ncu_target.py 4096 1 > gpurun_out/plain_solve4096.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:se3mpc_solve_kernel -s 1 -c 1 -o gpurun_out/r2_solve_4096 \
    python profiles/ncu_target.py 4096 1 > gpurun_out/ncu_solve4096.log 2>&1
python profiles/lone_probe.py > gpurun_out/r2_lone_final.txt 2>&1
cat gpurun_out/r2_final_gputests.txt; tail -2 gpurun_out/plain_solve4096.log; tail -c 600 gpurun_out/r2_bench.err
