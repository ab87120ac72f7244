set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_deddqn_run.py tests/test_gpu_sweep.py tests/test_gpu_tester.py -q -m gpu -p no:cacheprovider ) 2>&1 | tail -3
( timeout 900 python bench.py --workload deddqn --steps 3 --warmup 2 ) > gpurun_out/r2s_bench_n1_deddqn.json 2> gpurun_out/r2s.err; tail -1 gpurun_out/r2s_bench_n1_deddqn.json | cut -c1-300; tail -2 gpurun_out/r2s.err
