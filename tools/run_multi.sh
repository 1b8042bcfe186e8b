#!/bin/bash
# usage: tools/run_multi.sh N   - cfg-3 (1024 x 10 s, strong scaling) and cfg-5 (16 x 4 s + loss per rank) on N GPUs of this box
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
if [ "$N" = "1" ]; then TR="python"; fi
$TR bench.py --gpus $N --steps 3 --warmup 3 --seconds 10 --total-utterances 1024 --micro-batch 32 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_cfg3_${N}gpu.log 2>&1
tail -1 gpurun_out/r02_cfg3_${N}gpu.log | cut -c1-400
$TR bench.py --gpus $N --steps 5 --warmup 3 --batch 16 --loss --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_cfg5_${N}gpu.log 2>&1
tail -1 gpurun_out/r02_cfg5_${N}gpu.log | cut -c1-400
