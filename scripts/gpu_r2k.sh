set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_eval.py tests/test_gpu_random_search.py tests/test_gpu_tester.py tests/test_gpu_lde.py -q -m gpu -p no:cacheprovider ) 2>&1 | tail -4
timeout 600 python scripts/kernel_rooflines.py --only eval --fids 1,5,8,20 > gpurun_out/r2k_eval_rooflines.txt 2>&1; cat gpurun_out/r2k_eval_rooflines.txt
