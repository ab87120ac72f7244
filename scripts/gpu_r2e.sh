set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:deddqn_run -s 1 -c 1 -o gpurun_out/r2e_ncu_ddrun -f python scripts/phase_profile_deddqn_run.py 256 40 > gpurun_out/r2e_ncu.log 2>&1; tail -3 gpurun_out/r2e_ncu.log
