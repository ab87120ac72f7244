set -x
python -m pytest tests/test_gpu_rlepso.py tests/test_gpu_sweep.py -x -q -m gpu 2>&1 | tail -15
python scripts/per_function.py --suit bbob --dim 30 --algos rlepso > gpurun_out/per_function_rlepso_r1d.txt 2>&1; tail -26 gpurun_out/per_function_rlepso_r1d.txt
( timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_r1d_rlepso.log 2>&1; tail -2 gpurun_out/bench_r1d_rlepso.log | cut -c1-1800
