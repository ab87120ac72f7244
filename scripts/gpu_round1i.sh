set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_rlepso.py tests/test_gpu_sweep.py -x -q -m gpu ) 2>&1 | tail -15
MBX_LIB=$PWD/metabox_b200/csrc/libmbx_prof.so timeout 600 python scripts/phase_profile.py --fids all,1,21 > gpurun_out/phase_r1i.txt 2>&1
MBX_LIB=$PWD/metabox_b200/csrc/libmbx_prof.so timeout 600 python scripts/phase_profile.py --warm 350 --fids all,1,21 >> gpurun_out/phase_r1i.txt 2>&1
cat gpurun_out/phase_r1i.txt
( time timeout 1200 python bench.py --no-cpu-baseline ) > gpurun_out/bench_r1i.log 2>&1; grep -o '"value": [0-9.]*\|"avg_launch_ms": [0-9.]*' gpurun_out/bench_r1i.log
