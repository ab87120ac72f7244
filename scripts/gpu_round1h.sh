set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -15
timeout 600 python scripts/per_function.py --suit bbob --dim 30 --algos eval,rlepso > gpurun_out/per_function_r1h.txt 2>&1; cat gpurun_out/per_function_r1h.txt | sed -n 21,26p
( time timeout 1200 python bench.py --no-cpu-baseline ) > gpurun_out/bench_r1h.log 2>&1; grep -o '"value": [0-9.]*\|"avg_launch_ms": [0-9.]*' gpurun_out/bench_r1h.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rlepso_step -s 100 -c 1 -o gpurun_out/prof_rlepso_r1h python bench.py --steps 1 --warmup 0 --runs 64 --no-cpu-baseline > gpurun_out/ncu_r1h.log 2>&1; tail -3 gpurun_out/ncu_r1h.log
