set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider ) > gpurun_out/r2g_gpu_suite.log 2>&1; tail -15 gpurun_out/r2g_gpu_suite.log
( timeout 600 python bench.py --workload deddqn --steps 2 --warmup 1 ) > gpurun_out/r2g_bench_n1_deddqn.json 2> gpurun_out/r2g_deddqn.err; tail -1 gpurun_out/r2g_bench_n1_deddqn.json | cut -c1-300; tail -3 gpurun_out/r2g_deddqn.err
( timeout 600 python bench.py --workload protein --steps 5 --warmup 3 ) > gpurun_out/r2g_bench_n1_protein.json 2> gpurun_out/r2g_protein.err; tail -1 gpurun_out/r2g_bench_n1_protein.json | cut -c1-300; tail -3 gpurun_out/r2g_protein.err
( timeout 600 python bench.py --workload lde-train --steps 10 --warmup 3 ) > gpurun_out/r2g_bench_n1_lde_train.json 2> gpurun_out/r2g_lde_train.err; tail -1 gpurun_out/r2g_bench_n1_lde_train.json | cut -c1-300; tail -3 gpurun_out/r2g_lde_train.err
( timeout 600 python bench.py --workload gleet --steps 3 --warmup 3 ) > gpurun_out/r2g_bench_n1_gleet.json 2> gpurun_out/r2g_gleet.err; tail -1 gpurun_out/r2g_bench_n1_gleet.json | cut -c1-300; tail -3 gpurun_out/r2g_gleet.err
