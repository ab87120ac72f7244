set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -6
( time timeout 1200 python bench.py ) > gpurun_out/r1e_bench_n1_rlepso.log 2>&1; tail -1 gpurun_out/r1e_bench_n1_rlepso.log | cut -c1-300
( time timeout 1200 python bench.py --workload lde --steps 3 ) > gpurun_out/r1e_bench_n1_lde.log 2>&1; tail -1 gpurun_out/r1e_bench_n1_lde.log | cut -c1-200
( time timeout 1500 python bench.py --workload deddqn --steps 2 ) > gpurun_out/r1e_bench_n1_deddqn.log 2>&1; tail -1 gpurun_out/r1e_bench_n1_deddqn.log | cut -c1-200
timeout 600 python scripts/per_function.py --suit bbob --dim 30 --algos eval,rlepso,lde,gleet,deddqn > gpurun_out/r1e_per_function_bbob_d30.txt 2>&1; tail -26 gpurun_out/r1e_per_function_bbob_d30.txt
MBX_LIB=$PWD/metabox_b200/csrc/libmbx_prof.so timeout 600 python scripts/phase_profile.py --fids all,1,15,21 > gpurun_out/r1e_phase_profile.txt 2>&1; cat gpurun_out/r1e_phase_profile.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1e_launches.csv python bench.py --steps 1 --warmup 0 --runs 64 --no-cpu-baseline --stripes 1 > gpurun_out/r1e_launches.log 2>&1; tail -2 gpurun_out/r1e_launches.log | cut -c1-200
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:rlepso_step -s 60 -c 1 -o gpurun_out/r1e_ncu_rlepso python bench.py --steps 1 --warmup 0 --runs 256 --no-cpu-baseline --stripes 1 > gpurun_out/r1e_ncu_rlepso.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:lde_step -s 60 -c 1 -o gpurun_out/r1e_ncu_lde python bench.py --workload lde --steps 1 --warmup 0 --runs 256 --no-cpu-baseline --stripes 1 > gpurun_out/r1e_ncu_lde.log 2>&1
ls -la gpurun_out/*.ncu-rep
