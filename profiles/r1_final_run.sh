set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v11c.json 2> gpurun_out/bench_v11c.err; cat gpurun_out/bench_v11c.json; tail -1 gpurun_out/bench_v11c.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_v11c_ref.json 2>/dev/null; cat gpurun_out/bench_v11c_ref.json
timeout 300 python profiles/parity_census.py > gpurun_out/parity_census_v11c.txt 2>&1; tail -5 gpurun_out/parity_census_v11c.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v11c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_v11c.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:se3mpc_solve -s 1 -c 1 -o gpurun_out/v11c_full -f python profiles/ncu_target.py 4096 1 > gpurun_out/ncu_full_v11c.log 2>&1; tail -2 gpurun_out/ncu_full_v11c.log
