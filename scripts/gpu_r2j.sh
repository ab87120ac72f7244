set -x
mkdir -p gpurun_out
timeout 600 python scripts/kernel_rooflines.py > gpurun_out/r2j_kernel_rooflines.txt 2>&1; cat gpurun_out/r2j_kernel_rooflines.txt
( timeout 300 python -m pytest tests/test_gpu_eval.py -q -m gpu -p no:cacheprovider ) 2>&1 | tail -3
