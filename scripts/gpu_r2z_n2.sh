# two-GPU sanity of the final round-2 build: headline + the LDE train step through torchrun (gpurun --gpus 2)
set -x
mkdir -p gpurun_out
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r2z_bench_n2_rlepso.json 2> gpurun_out/r2z_n2.err; tail -1 gpurun_out/r2z_bench_n2_rlepso.json | cut -c1-260; tail -2 gpurun_out/r2z_n2.err
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --workload lde-train --steps 10 --warmup 3 ) > gpurun_out/r2z_bench_n2_lde_train.json 2> gpurun_out/r2z_n2b.err; tail -1 gpurun_out/r2z_bench_n2_lde_train.json | cut -c1-260; tail -2 gpurun_out/r2z_n2b.err
