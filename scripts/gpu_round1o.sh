mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -15
for st in 1 2 3; do ( timeout 900 python bench.py --no-cpu-baseline --stripes $st ) > gpurun_out/bench_r1o_s$st.log 2>&1; echo stripes $st; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_r1o_s$st.log; tail -3 gpurun_out/bench_r1o_s$st.log | grep -i error; done
( timeout 900 python bench.py --workload lde --no-cpu-baseline --steps 3 --warmup 3 ) > gpurun_out/bench_r1o_lde.log 2>&1; grep -o '"value": [0-9.]*' gpurun_out/bench_r1o_lde.log
