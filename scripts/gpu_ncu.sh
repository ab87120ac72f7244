set -x
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:rlepso_step -s 60 -c 1 -o gpurun_out/prof_rlepso_r1j python bench.py --steps 1 --warmup 0 --runs 256 --no-cpu-baseline > gpurun_out/ncu_r1j.log 2>&1; tail -2 gpurun_out/ncu_r1j.log | cut -c1-300
