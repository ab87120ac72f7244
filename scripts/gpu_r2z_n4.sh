# four-GPU line of the final round-2 build (gpurun --gpus 4)
set -x
mkdir -p gpurun_out
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 5 --warmup 3 ) > gpurun_out/r2z_bench_n4_rlepso.json 2> gpurun_out/r2z_n4.err; tail -1 gpurun_out/r2z_bench_n4_rlepso.json | cut -c1-200; tail -2 gpurun_out/r2z_n4.err
