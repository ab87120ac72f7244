set -x
nvidia-smi -L
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time timeout 900 python bench.py --steps 2 --warmup 1 ) > gpurun_out/bench_r1a.log 2>&1
tail -5 gpurun_out/bench_r1a.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 1 --warmup 0 --runs 64 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rlepso_step -s 100 -c 2 -o gpurun_out/prof_rlepso_r1a python bench.py --steps 1 --warmup 0 --runs 64 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
