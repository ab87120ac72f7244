set -x
python -m pytest tests/test_gpu_sweep.py -x -q -m gpu 2>&1 | tail -30
( time timeout 900 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench_r1b_rlepso.log 2>&1; tail -3 gpurun_out/bench_r1b_rlepso.log
( time timeout 900 python bench.py --steps 3 --warmup 3 --workload lde ) > gpurun_out/bench_r1b_lde.log 2>&1; tail -3 gpurun_out/bench_r1b_lde.log
( time timeout 1200 python bench.py --steps 2 --warmup 3 --workload deddqn ) > gpurun_out/bench_r1b_deddqn.log 2>&1; tail -3 gpurun_out/bench_r1b_deddqn.log
