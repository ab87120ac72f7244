set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -x -q -m gpu --timeout 200 ) 2>&1 | tail -6
( timeout 300 python bench.py ) > gpurun_out/r1g_bench_n1_rlepso.log 2>&1; tail -1 gpurun_out/r1g_bench_n1_rlepso.log | cut -c1-150
( timeout 300 python bench.py --workload lde --steps 3 ) > gpurun_out/r1g_bench_n1_lde.log 2>&1; tail -1 gpurun_out/r1g_bench_n1_lde.log | cut -c1-150
( timeout 300 python bench.py --workload gleet --steps 3 ) > gpurun_out/r1g_bench_n1_gleet.log 2>&1; tail -1 gpurun_out/r1g_bench_n1_gleet.log | cut -c1-150
( timeout 400 python bench.py --workload deddqn --steps 2 ) > gpurun_out/r1g_bench_n1_deddqn.log 2>&1; tail -1 gpurun_out/r1g_bench_n1_deddqn.log | cut -c1-150
( timeout 300 python bench.py --workload deddqn --steps 1 --warmup 1 --stripes 12 --no-cpu-baseline ) 2>&1 | tail -1 | cut -c1-150
for algo in gleet rlepso-protein deddqn rl_pso; do
  k=$(echo $algo | sed 's/rlepso-protein/rlepso_step/; s/gleet/gleet_step/; s/deddqn/deddqn_step/; s/rl_pso/rlpso_step/')
  timeout 200 ncu --set full --clock-control none -k regex:$k -s 20 -c 1 -o gpurun_out/r1g_ncu_$algo python scripts/ncu_target.py --algo $algo > gpurun_out/r1g_ncu_$algo.log 2>&1; tail -1 gpurun_out/r1g_ncu_$algo.log
done
ls -la gpurun_out/*.ncu-rep | tail -5
