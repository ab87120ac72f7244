# Round 2, call B: new parity tests (benchmark shapes, persistent DE-DDQN), then the DE-DDQN bench with the persistent kernel.
set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_deddqn_run.py tests/test_gpu_deddqn.py tests/test_gpu_benchmark_shapes.py tests/test_gpu_z_policies.py -q -m gpu --timeout 600 -p no:cacheprovider ) > gpurun_out/r2b_tests.log 2>&1; tail -40 gpurun_out/r2b_tests.log
( timeout 600 python bench.py --workload deddqn --steps 2 --warmup 1 --no-cpu-baseline ) > gpurun_out/r2b_bench_n1_deddqn_persistent.json 2> gpurun_out/r2b_deddqn.err; tail -1 gpurun_out/r2b_bench_n1_deddqn_persistent.json | cut -c1-400; tail -5 gpurun_out/r2b_deddqn.err
( timeout 600 python bench.py --workload deddqn --steps 2 --warmup 1 --no-cpu-baseline --stripes 1 ) > gpurun_out/r2b_bench_n1_deddqn_persistent_s1.json 2> gpurun_out/r2b_deddqn_s1.err; tail -1 gpurun_out/r2b_bench_n1_deddqn_persistent_s1.json | cut -c1-400; tail -5 gpurun_out/r2b_deddqn_s1.err
