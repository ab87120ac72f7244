set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python scripts/per_function.py --suit bbob --dim 30 > gpurun_out/per_function_bbob_d30_r1c.txt 2>&1; tail -30 gpurun_out/per_function_bbob_d30_r1c.txt
