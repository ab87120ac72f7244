set -x
mkdir -p gpurun_out
( time timeout 900 python bench.py --workload lde --no-cpu-baseline --steps 3 --warmup 3 ) > gpurun_out/bench_r1k_lde.log 2>&1; grep -o '"value": [0-9.]*\|"avg_launch_ms": [0-9.]*\|"kernel_share_of_step": [0-9.]*' gpurun_out/bench_r1k_lde.log
( time timeout 900 python bench.py --workload deddqn --no-cpu-baseline --steps 1 --warmup 1 ) > gpurun_out/bench_r1k_deddqn.log 2>&1; grep -o '"value": [0-9.]*\|"avg_launch_ms": [0-9.]*\|"kernel_share_of_step": [0-9.]*' gpurun_out/bench_r1k_deddqn.log
timeout 600 python scripts/per_function.py --suit bbob --dim 30 --algos lde,deddqn > gpurun_out/per_function_r1k.txt 2>&1; cat gpurun_out/per_function_r1k.txt
