set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( time timeout 900 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -8
MBX_LIB=$PWD/metabox_b200/csrc/libmbx_prof.so timeout 600 python scripts/phase_profile.py > gpurun_out/phase_r1e.txt 2>&1; cat gpurun_out/phase_r1e.txt
( time timeout 1200 python bench.py ) > gpurun_out/bench_r1e_default.log 2>&1; tail -4 gpurun_out/bench_r1e_default.log | cut -c1-2500
