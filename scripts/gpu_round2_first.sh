# First GPU call of round 2: validate what round 1 wrote after its GPU minutes were spent (fused policy kernels, GLEET trainer), then measure.
set -x
mkdir -p gpurun_out
# --runxfail: the pending markers are ignored, every failure is shown with its traceback (no -x: one bad test must not hide the others)
( time timeout 900 python -m pytest tests/test_gpu_z_policies.py -q -m gpu --runxfail --timeout 240 ) > gpurun_out/r2a_pending_tests.log 2>&1; tail -40 gpurun_out/r2a_pending_tests.log
timeout 200 python scripts/policy_timing.py > gpurun_out/r2a_policy_timing.txt 2>&1; cat gpurun_out/r2a_policy_timing.txt
( timeout 300 python bench.py --workload gleet --gleet-policy fused --steps 3 --no-cpu-baseline ) > gpurun_out/r2a_bench_n1_gleet_fused.json 2> gpurun_out/r2a_gleet_fused.err; tail -1 gpurun_out/r2a_bench_n1_gleet_fused.json | cut -c1-200
( timeout 500 python bench.py --workload deddqn --fused-policy --steps 2 --no-cpu-baseline ) > gpurun_out/r2a_bench_n1_deddqn_fused.json 2> gpurun_out/r2a_deddqn_fused.err; tail -1 gpurun_out/r2a_bench_n1_deddqn_fused.json | cut -c1-200
( timeout 300 python bench.py --fused-policy --steps 3 --no-cpu-baseline ) > gpurun_out/r2a_bench_n1_rlepso_fused.json 2> gpurun_out/r2a_rlepso_fused.err; tail -1 gpurun_out/r2a_bench_n1_rlepso_fused.json | cut -c1-200
( timeout 300 python bench.py --steps-per-graph 4 --steps 3 --no-cpu-baseline ) > gpurun_out/r2a_bench_n1_rlepso_spg4.json 2> gpurun_out/r2a_rlepso_spg4.err; tail -1 gpurun_out/r2a_bench_n1_rlepso_spg4.json | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gleet_policy -s 3 -c 1 -o gpurun_out/r2a_ncu_gleet_policy -f python scripts/policy_timing.py --only gleet --reps 4 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:deddqn_policy -s 3 -c 1 -o gpurun_out/r2a_ncu_deddqn_policy -f python scripts/policy_timing.py --only deddqn --reps 4 > /dev/null 2>&1
ls -la gpurun_out | tail -8
timeout 600 python scripts/train_timing.py > gpurun_out/r2a_train_timing.txt 2>&1; cat gpurun_out/r2a_train_timing.txt
timeout 600 python scripts/composed_timing.py > gpurun_out/r2a_composed_timing.txt 2>&1; cat gpurun_out/r2a_composed_timing.txt
# A/B of the packed-f32x2 variant of the GLEET policy kernel (build it first in the build container:
#   python -m metabox_b200.build --variant glp2 -DGLP_PACKED_F32X2 ): same timing script, selected through MBX_LIB
if [ -f metabox_b200/csrc/libmbx_glp2.so ]; then MBX_LIB=$PWD/metabox_b200/csrc/libmbx_glp2.so timeout 200 python scripts/policy_timing.py --only gleet > gpurun_out/r2a_policy_timing_glp2.txt 2>&1; cat gpurun_out/r2a_policy_timing_glp2.txt; fi
