set -x
mkdir -p gpurun_out
( time timeout 500 python -m pytest tests/test_gpu_gleet.py tests/test_gpu_rl_pso.py tests/test_gpu_deddqn.py tests/test_gpu_eval.py tests/test_gpu_tester.py tests/test_gpu_configs.py tests/test_gpu_edge_sizes.py -x -q -m gpu --timeout 200 ) 2>&1 | tail -6
for algo in gleet rl_pso deddqn; do timeout 120 python scripts/ncu_target.py --algo $algo --steps 60 | tail -1; done
( timeout 300 python bench.py --workload gleet --steps 3 --no-cpu-baseline ) 2>&1 | tail -1 | cut -c1-200
( timeout 300 python bench.py --workload deddqn --steps 2 --no-cpu-baseline ) 2>&1 | tail -1 | cut -c1-200
