# multi-GPU lines of the fixed-size configs (strong scaling): run with gpurun --gpus N
set -x
N=$1
mkdir -p gpurun_out
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload lde-train --steps 20 --warmup 5 ) > gpurun_out/r2h_bench_n${N}_lde_train.json 2> gpurun_out/r2h_lde_train_n${N}.err; tail -1 gpurun_out/r2h_bench_n${N}_lde_train.json | cut -c1-300; tail -3 gpurun_out/r2h_lde_train_n${N}.err
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --workload protein --steps 10 --warmup 3 ) > gpurun_out/r2h_bench_n${N}_protein.json 2> gpurun_out/r2h_protein_n${N}.err; tail -1 gpurun_out/r2h_bench_n${N}_protein.json | cut -c1-300; tail -3 gpurun_out/r2h_protein_n${N}.err
