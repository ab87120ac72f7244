set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -25
timeout 600 python scripts/per_function.py --suit bbob --dim 30 --algos eval,rlepso,lde,deddqn > gpurun_out/per_function_r1l.txt 2>&1; cat gpurun_out/per_function_r1l.txt
( time timeout 1200 python bench.py --no-cpu-baseline ) > gpurun_out/bench_r1l.log 2>&1; grep -o '"value": [0-9.]*\|"avg_launch_ms": [0-9.]*' gpurun_out/bench_r1l.log
