set -x
mkdir -p gpurun_out
run() { ( timeout 600 python bench.py --workload deddqn --steps 2 --warmup 1 --no-cpu-baseline $@ ) 2> gpurun_out/r2r.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['value']/1e6,2), 'M evals/s', round(d['ms_per_step'],1), 'ms')"; tail -2 gpurun_out/r2r.err; }
run --stripes 3 --launches-per-poll 16
run --stripes 3 --poll 1000 --launches-per-poll 8
run --stripes 3 --poll 250 --launches-per-poll 16
run --stripes 6 --launches-per-poll 8
run --stripes 5 --launches-per-poll 8
