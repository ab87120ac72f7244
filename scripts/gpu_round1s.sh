mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_deddqn.py tests/test_gpu_sweep.py -x -q -m gpu 2>&1 | tail -4
MBX_LIB=$PWD/metabox_b200/csrc/libmbx_prof.so timeout 300 python scripts/phase_profile_deddqn.py 2>&1 | tail -3
( timeout 1500 python bench.py --workload deddqn --no-cpu-baseline --steps 1 --warmup 1 ) > gpurun_out/bench_r1s_deddqn.log 2>&1; grep -o '"value": [0-9.]*\|"avg_launch_ms": [0-9.]*\|"kernel_share_of_step": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_r1s_deddqn.log
