# A/B of compile-time tuning knobs of the headline kernel (metabox_b200/csrc/mbx_rlepso.cu: MBX_FDR_UNROLL, MBX_RLEPSO_MIN_CTAS, MBX_RLEPSO_P2_CH).
#   step 1 (build container, no GPU):   bash scripts/variant_sweep.sh build     -> metabox_b200/csrc/libmbx_<name>.so (travel with gpurun)
#   step 2 (GPU box):                   bash scripts/variant_sweep.sh bench     -> gpurun_out/variants.txt
# Every variant must pass the bit-exact RLEPSO device tests before its number counts (the `bench` step runs them first).
set -e
VARIANTS=${VARIANTS:-"fdr1:-DMBX_FDR_UNROLL=1 fdr4:-DMBX_FDR_UNROLL=4 ctas3:-DMBX_RLEPSO_MIN_CTAS=3 fdr4ctas3:-DMBX_FDR_UNROLL=4,-DMBX_RLEPSO_MIN_CTAS=3 p2ch6:-DMBX_RLEPSO_P2_CH=6 p2ch24:-DMBX_RLEPSO_P2_CH=24"}
if [ "$1" = "build" ]; then
  for v in $VARIANTS; do
    name=${v%%:*}; defs=$(echo ${v#*:} | tr ',' ' ')
    python -m metabox_b200.build --variant $name $defs
    grep -h "rlepso_step_kernel" -A2 metabox_b200/csrc/build_$name/mbx_rlepso.o.log | grep -E "registers|spill" | head -4
  done
else
  mkdir -p gpurun_out
  : > gpurun_out/variants.txt
  for lib in default $(for v in $VARIANTS; do echo ${v%%:*}; done); do
    if [ $lib = default ]; then unset MBX_LIB; else export MBX_LIB=$PWD/metabox_b200/csrc/libmbx_$lib.so; fi
    if timeout 600 python -m pytest tests/test_gpu_rlepso.py -x -q -m gpu > gpurun_out/variant_$lib.tests 2>&1; then ok=green; else ok=RED; fi
    line=$(timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), 'M evals/s, step kernel', round(d['roofline']['avg_launch_ms'],4), 'ms')")
    echo "$lib: tests $ok, $line" | tee -a gpurun_out/variants.txt
  done
fi
