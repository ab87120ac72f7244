# round 2, late: full device suite + the single-GPU bench lines after the FDR key scan / TEST instantiations / u01 change
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/r2z_gpu_suite.log
for w in rlepso lde gleet deddqn protein lde-train; do
  ( timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline ) 2>gpurun_out/r2z_$w.err | tail -1 > gpurun_out/r2z_bench_n1_$w.json
  python -c "import json,sys; d=json.loads(open('gpurun_out/r2z_bench_n1_$w.json').read()); print('$w', round(d['value']/1e6,2), 'M evals/s; e2e', round(d['e2e']['value']/1e6,2), '; kernel ms', d['roofline'].get('avg_launch_ms'))"
done
