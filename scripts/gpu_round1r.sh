mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_deddqn.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python scripts/per_function.py --suit bbob --dim 30 --algos deddqn 2>&1 | sed -n 1,6p
( timeout 1500 python bench.py --workload deddqn --no-cpu-baseline --steps 1 --warmup 1 ) > gpurun_out/bench_r1r_deddqn.log 2>&1; grep -o '"value": [0-9.]*\|"avg_launch_ms": [0-9.]*\|"kernel_share_of_step": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_r1r_deddqn.log
