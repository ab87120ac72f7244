set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_gleet.py tests/test_gpu_z_policies.py tests/test_gpu_edge_sizes.py tests/test_gpu_trainer.py -q -m gpu -p no:cacheprovider ) 2>&1 | tail -2
for i in 1 2; do ( timeout 300 python bench.py --workload gleet --steps 5 --warmup 3 --no-cpu-baseline ) 2>gpurun_out/r2w.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('gleet', round(d['value']/1e6,2), 'M evals/s', round(d['ms_per_step'],1), 'ms; step kernel', round(d['roofline']['avg_launch_ms'],4))"; done
