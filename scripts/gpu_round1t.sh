mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eval.py tests/test_gpu_rlepso.py tests/test_gpu_lde.py tests/test_gpu_deddqn.py tests/test_gpu_gleet.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -6
python scripts/config_timing.py 2>&1 | tail -3
