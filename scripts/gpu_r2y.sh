set -x
mkdir -p gpurun_out
timeout 300 python scripts/policy_timing.py --only gleet 2>&1 | tail -2
MBX_LIB=$PWD/metabox_b200/csrc/libmbx_glp2.so timeout 300 python scripts/policy_timing.py --only gleet 2>&1 | tail -2
MBX_LIB=$PWD/metabox_b200/csrc/libmbx_glp2.so timeout 600 python -m pytest tests/test_gpu_z_policies.py -q -m gpu -p no:cacheprovider -k gleet 2>&1 | tail -2
( MBX_LIB=$PWD/metabox_b200/csrc/libmbx_glp2.so timeout 300 python bench.py --workload gleet --steps 5 --warmup 3 --no-cpu-baseline ) 2>gpurun_out/r2y.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('gleet glp2', round(d['value']/1e6,2), 'M evals/s')"
