# ncu --set full captures of the kernels that had none (J2): one launch each; summarised ON the box (the reports together exceed the 64 MiB
# that come back), only the text summaries return
set -x
mkdir -p gpurun_out /tmp/ncu
cap() {  # name, kernel regex, skip, script args
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o /tmp/ncu/$1 -f python scripts/kernel_rooflines.py --reps 2 $4 > /tmp/ncu/$1.log 2>&1
  ( echo "# ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 python scripts/kernel_rooflines.py --reps 2 $4"; python scripts/ncu_summary.py /tmp/ncu/$1.ncu-rep ) > gpurun_out/r2l_ncu_$1.txt 2>&1
  tail -2 gpurun_out/r2l_ncu_$1.txt
}
cap eval_stream_f8 eval_stream 3 "--only eval --fids 8"
cap eval_f1 'eval_kernel' 3 "--only eval --fids 1"
cap eval_f15 'eval_kernel' 3 "--only eval --fids 15"
cap random_search random_search 3 "--only rs"
cap rl_pso_step rlpso_step 3 "--only rl_pso"
cap qlpso_step qlpso_step 3 "--only qlpso"
cap landscape landscape 3 "--only landscape"
cap walk progressive_walk 3 "--only walk"
cap gleet_policy gleet_policy 3 "--only policies"
cap deddqn_policy deddqn_policy 3 "--only policies"
cap rlepso_policy rlepso_policy 3 "--only policies"
cap rl_pso_policy rl_pso_policy 3 "--only policies"
cap qlpso_policy qlpso_policy 3 "--only policies"
timeout 600 python scripts/kernel_rooflines.py > gpurun_out/r2l_kernel_rooflines.txt 2>&1
( timeout 900 python -m pytest tests/test_gpu_eval.py tests/test_gpu_benchmark_shapes.py tests/test_gpu_deddqn_run.py tests/test_gpu_rlepso.py -q -m gpu -p no:cacheprovider -k "dim_20 or 20 or eval" ) 2>&1 | tail -4
du -sh gpurun_out
