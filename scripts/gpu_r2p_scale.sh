# weak-scaling check of the headline at N GPUs (the driver's command line), both arms
set -x
N=$1
mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --impl reference --gpus $N --steps 3 --warmup 1 ) > gpurun_out/r2p_bench_reference_n${N}.json 2> gpurun_out/r2p_ref_n${N}.err; tail -1 gpurun_out/r2p_bench_reference_n${N}.json | cut -c1-250
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r2p_bench_n${N}_rlepso.json 2> gpurun_out/r2p_n${N}.err; tail -1 gpurun_out/r2p_bench_n${N}_rlepso.json | cut -c1-250; tail -3 gpurun_out/r2p_n${N}.err
nproc
