set -x
mkdir -p gpurun_out
run() { ( timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $@ ) 2> gpurun_out/r2t.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['value']/1e6,2), 'M evals/s', round(d['ms_per_step'],1), 'ms')"; tail -2 gpurun_out/r2t.err; }
run --poll 64
run --poll 200
run --poll 700
run --poll 200 --stripes 4
run --poll 200 --stripes 2
run --workload gleet --poll 200
run --workload lde --poll 200
run --workload lde
