mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_tester.py -x -q -m gpu ) 2>&1 | tail -25
( timeout 900 python bench.py --no-cpu-baseline --stripes 4 ) > gpurun_out/bench_r1p_s4.log 2>&1; echo stripes 4; grep -o '"value": [0-9.]*' gpurun_out/bench_r1p_s4.log
( time timeout 1500 python bench.py --workload deddqn --no-cpu-baseline --steps 1 --warmup 1 ) > gpurun_out/bench_r1p_deddqn.log 2>&1; grep -o '"value": [0-9.]*\|"avg_launch_ms": [0-9.]*\|"kernel_share_of_step": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_r1p_deddqn.log; tail -2 gpurun_out/bench_r1p_deddqn.log | cut -c1-300
