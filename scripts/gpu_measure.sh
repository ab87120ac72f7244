set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -6
( time timeout 1200 python bench.py ) > gpurun_out/r1g_bench_n1_rlepso.log 2>&1; tail -1 gpurun_out/r1g_bench_n1_rlepso.log | cut -c1-200
( time timeout 1200 python bench.py --impl reference ) > gpurun_out/r1g_bench_reference.log 2>&1; tail -1 gpurun_out/r1g_bench_reference.log | cut -c1-200
( time timeout 1200 python bench.py --workload lde --steps 3 ) > gpurun_out/r1g_bench_n1_lde.log 2>&1; tail -1 gpurun_out/r1g_bench_n1_lde.log | cut -c1-200
( time timeout 1500 python bench.py --workload deddqn --steps 2 ) > gpurun_out/r1g_bench_n1_deddqn.log 2>&1; tail -1 gpurun_out/r1g_bench_n1_deddqn.log | cut -c1-200
timeout 600 python scripts/per_function.py --suit bbob --dim 30 --algos eval,rlepso,lde,gleet,deddqn > gpurun_out/r1g_per_function_bbob_d30.txt 2>&1
timeout 600 python scripts/config_timing.py > gpurun_out/r1g_config_timing.txt 2>&1; cat gpurun_out/r1g_config_timing.txt
python __graft_entry__.py smoke 2>&1 | tail -1
