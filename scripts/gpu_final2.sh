set -x
mkdir -p gpurun_out
( time timeout 700 python -m pytest tests -x -q -m gpu --timeout 200 ) 2>&1 | tail -6
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( timeout 300 python bench.py --workload gleet --steps 3 ) > gpurun_out/r1h_bench_n1_gleet.log 2>&1; tail -1 gpurun_out/r1h_bench_n1_gleet.log | cut -c1-150
( timeout 400 python bench.py --workload deddqn --steps 2 ) > gpurun_out/r1h_bench_n1_deddqn.log 2>&1; tail -1 gpurun_out/r1h_bench_n1_deddqn.log | cut -c1-150
