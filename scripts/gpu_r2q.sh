set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_deddqn_run.py tests/test_gpu_sweep.py -q -m gpu -p no:cacheprovider ) 2>&1 | tail -3
for s in 1 2 3 4; do
( timeout 600 python bench.py --workload deddqn --steps 2 --warmup 1 --no-cpu-baseline --stripes $s ) > gpurun_out/r2q_bench_n1_deddqn_s$s.json 2> gpurun_out/r2q_s$s.err; tail -1 gpurun_out/r2q_bench_n1_deddqn_s$s.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('stripes $s', round(d['value']/1e6,2), 'M evals/s', round(d['ms_per_step'],1), 'ms')"; tail -2 gpurun_out/r2q_s$s.err
done
timeout 900 python scripts/train_timing.py --only gleet > gpurun_out/r2q_train_timing_gleet.txt 2>&1; tail -3 gpurun_out/r2q_train_timing_gleet.txt
