set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -15
timeout 600 python scripts/per_function.py --suit bbob --dim 30 --algos eval,rlepso > gpurun_out/per_function_r1g.txt 2>&1; cat gpurun_out/per_function_r1g.txt
( time timeout 1200 python bench.py --no-cpu-baseline ) > gpurun_out/bench_r1g.log 2>&1; tail -4 gpurun_out/bench_r1g.log | cut -c1-400
