set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_deddqn_run.py tests/test_gpu_deddqn.py tests/test_gpu_benchmark_shapes.py tests/test_gpu_sweep.py tests/test_gpu_configs.py -q -m gpu --timeout 600 -p no:cacheprovider -k "deddqn or c3" ) > gpurun_out/r2d_tests.log 2>&1; tail -30 gpurun_out/r2d_tests.log
timeout 300 python scripts/phase_profile_deddqn_run.py > gpurun_out/r2d_ddrun_timing.txt 2>&1; cat gpurun_out/r2d_ddrun_timing.txt
MBX_LIB=$PWD/metabox_b200/csrc/libmbx_prof.so timeout 300 python scripts/phase_profile_deddqn_run.py > gpurun_out/r2d_ddrun_phases.txt 2>&1; cat gpurun_out/r2d_ddrun_phases.txt
timeout 300 python -m pytest tests/test_gpu_z_policies.py -q -m gpu -k rl_hpsde -p no:cacheprovider 2>&1 | tail -3
