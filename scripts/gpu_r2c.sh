set -x
mkdir -p gpurun_out
timeout 300 python scripts/phase_profile_deddqn_run.py > gpurun_out/r2c_ddrun_timing.txt 2>&1; cat gpurun_out/r2c_ddrun_timing.txt
MBX_LIB=$PWD/metabox_b200/csrc/libmbx_prof.so timeout 300 python scripts/phase_profile_deddqn_run.py > gpurun_out/r2c_ddrun_phases.txt 2>&1; cat gpurun_out/r2c_ddrun_phases.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:deddqn_run -s 1 -c 1 -o gpurun_out/r2c_ncu_ddrun -f python scripts/phase_profile_deddqn_run.py 256 40 > gpurun_out/r2c_ncu.log 2>&1; tail -3 gpurun_out/r2c_ncu.log
timeout 300 python -m pytest tests/test_gpu_z_policies.py -q -m gpu -k rl_hpsde -p no:cacheprovider 2>&1 | tail -3
