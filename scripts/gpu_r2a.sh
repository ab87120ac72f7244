# Round 2, call A: the whole GPU suite without -x (every failure visible), then timing of the fused policies and bench variants.
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider ) > gpurun_out/r2a_gpu_suite.log 2>&1; tail -60 gpurun_out/r2a_gpu_suite.log
timeout 200 python scripts/policy_timing.py > gpurun_out/r2a_policy_timing.txt 2>&1; cat gpurun_out/r2a_policy_timing.txt
( timeout 300 python bench.py --workload gleet --gleet-policy fused --steps 3 --no-cpu-baseline ) > gpurun_out/r2a_bench_n1_gleet_fused.json 2> gpurun_out/r2a_gleet_fused.err; tail -1 gpurun_out/r2a_bench_n1_gleet_fused.json | cut -c1-300; tail -5 gpurun_out/r2a_gleet_fused.err
( timeout 500 python bench.py --workload deddqn --fused-policy --steps 2 --no-cpu-baseline ) > gpurun_out/r2a_bench_n1_deddqn_fused.json 2> gpurun_out/r2a_deddqn_fused.err; tail -1 gpurun_out/r2a_bench_n1_deddqn_fused.json | cut -c1-300; tail -5 gpurun_out/r2a_deddqn_fused.err
( timeout 300 python bench.py --fused-policy --steps 3 --no-cpu-baseline ) > gpurun_out/r2a_bench_n1_rlepso_fused.json 2> gpurun_out/r2a_rlepso_fused.err; tail -1 gpurun_out/r2a_bench_n1_rlepso_fused.json | cut -c1-300; tail -5 gpurun_out/r2a_rlepso_fused.err
( timeout 300 python bench.py --steps 3 --no-cpu-baseline ) > gpurun_out/r2a_bench_n1_rlepso.json 2> gpurun_out/r2a_rlepso.err; tail -1 gpurun_out/r2a_bench_n1_rlepso.json | cut -c1-300
timeout 600 python scripts/train_timing.py > gpurun_out/r2a_train_timing.txt 2>&1; cat gpurun_out/r2a_train_timing.txt
