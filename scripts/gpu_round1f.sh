set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_rlepso.py tests/test_gpu_sweep.py -x -q -m gpu ) 2>&1 | tail -8
export MBX_LIB=$PWD/metabox_b200/csrc/libmbx_prof.so
timeout 600 python scripts/phase_profile.py > gpurun_out/phase_r1f.txt 2>&1
timeout 600 python scripts/phase_profile.py --warm 350 --fids all,1,15,21 >> gpurun_out/phase_r1f.txt 2>&1
cat gpurun_out/phase_r1f.txt
unset MBX_LIB
( time timeout 1200 python bench.py --no-cpu-baseline ) > gpurun_out/bench_r1f.log 2>&1; tail -4 gpurun_out/bench_r1f.log | cut -c1-1500
