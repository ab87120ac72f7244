mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_lde.py tests/test_gpu_sweep.py -x -q -m gpu ) 2>&1 | tail -15
( time timeout 900 python bench.py --workload lde --no-cpu-baseline --steps 3 --warmup 3 ) > gpurun_out/bench_r1n_lde.log 2>&1; grep -o '"value": [0-9.]*\|"avg_launch_ms": [0-9.]*\|"kernel_share_of_step": [0-9.]*' gpurun_out/bench_r1n_lde.log
