set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_rlepso.py tests/test_gpu_lde.py tests/test_gpu_benchmark_shapes.py tests/test_gpu_edge_sizes.py tests/test_gpu_configs.py tests/test_gpu_trainer.py -q -m gpu -p no:cacheprovider ) > gpurun_out/r2u_tests.log 2>&1; tail -3 gpurun_out/r2u_tests.log
for w in rlepso lde; do ( timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline ) 2>gpurun_out/r2u.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', round(d['value']/1e6,2), 'M evals/s', round(d['ms_per_step'],1), 'ms; step kernel', round(d['roofline']['avg_launch_ms'],4))"; done
( timeout 300 python bench.py --workload lde-train --steps 20 --warmup 5 --no-cpu-baseline ) 2>gpurun_out/r2u.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lde-train', round(d['value']/1e6,2), 'M evals/s', round(d['ms_per_step'],2), 'ms; step kernel', round(d['roofline']['avg_launch_ms'],4))"
