#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -c 1500 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
echo "== reference arm under torchrun"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -c 600
