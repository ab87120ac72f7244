set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_deddqn_run.py -q -m gpu --timeout 600 -p no:cacheprovider ) 2>&1 | tail -3
timeout 300 python scripts/phase_profile_deddqn_run.py 2>&1 | tee gpurun_out/r2f_ddrun_timing.txt
