# dress rehearsal of the driver's round-end sequence on a fresh box
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2zz_gpu_suite.log 2>&1; tail -5 gpurun_out/r2zz_gpu_suite.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -4
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2zz_bench_reference_arm.json 2> gpurun_out/r2zz_ref.err; tail -1 gpurun_out/r2zz_bench_reference_arm.json | cut -c1-400; tail -4 gpurun_out/r2zz_ref.err
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2zz_bench_n1_rlepso.json 2> gpurun_out/r2zz_ours.err; tail -1 gpurun_out/r2zz_bench_n1_rlepso.json | cut -c1-300; tail -4 gpurun_out/r2zz_ours.err
nproc; grep -m1 "model name" /proc/cpuinfo
