set -x
mkdir -p gpurun_out
for lib in default fdr4 default fdr4; do
  if [ $lib = default ]; then unset MBX_LIB; else export MBX_LIB=$PWD/metabox_b200/csrc/libmbx_$lib.so; fi
  ( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) 2>gpurun_out/r2x.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', round(d['value']/1e6,2), 'M evals/s; step kernel', round(d['roofline']['avg_launch_ms'],4))"
done
MBX_LIB=$PWD/metabox_b200/csrc/libmbx_fdr4.so timeout 600 python -m pytest tests/test_gpu_rlepso.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2
