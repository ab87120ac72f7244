set -x
mkdir -p gpurun_out
( MBX_RLEPSO_SPLIT=1 timeout 900 python -m pytest tests/test_gpu_rlepso.py tests/test_gpu_benchmark_shapes.py tests/test_gpu_edge_sizes.py tests/test_gpu_sweep.py tests/test_gpu_configs.py -q -m gpu --timeout 600 -p no:cacheprovider -k "rlepso or c2 or c5 or striped" ) 2>&1 | tail -5
( MBX_RLEPSO_SPLIT=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2i_bench_n1_rlepso_fused.json 2>gpurun_out/r2i_a.err; tail -1 gpurun_out/r2i_bench_n1_rlepso_fused.json | cut -c1-200
( MBX_RLEPSO_SPLIT=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2i_bench_n1_rlepso_split.json 2>gpurun_out/r2i_b.err; tail -1 gpurun_out/r2i_bench_n1_rlepso_split.json | cut -c1-200; tail -3 gpurun_out/r2i_b.err
( MBX_RLEPSO_SPLIT=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --stripes 1 ) > gpurun_out/r2i_bench_n1_rlepso_split_s1.json 2>gpurun_out/r2i_c.err; tail -1 gpurun_out/r2i_bench_n1_rlepso_split_s1.json | cut -c1-200
( MBX_RLEPSO_SPLIT=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --stripes 1 ) > gpurun_out/r2i_bench_n1_rlepso_fused_s1.json 2>gpurun_out/r2i_d.err; tail -1 gpurun_out/r2i_bench_n1_rlepso_fused_s1.json | cut -c1-200
