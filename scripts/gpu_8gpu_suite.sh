#!/bin/bash
# 8 GPUs: sharded == whole frame; 8K bench with peer stores and with NCCL; replicated history; 16K; config 5 (64 x 1080p streams)
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
f() { grep -v "OMP_NUM\|^W1018\|^\*\*\*" | grep "AssertionError\|^{\|whole-frame check\|sharded_check\|Error" | head -6 | cut -c1-3000; }
{
echo "== sharded_check N=$N"; timeout 600 $TR --master-port 29511 scripts/sharded_check.py 2>&1 | f
echo "== 8K peer N=$N"; timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 96 --warmup 5 2>&1 | f
echo "== 8K nccl N=$N"; TAA_SHARDED_EXCHANGE=nccl timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 96 --warmup 5 --no-verify 2>&1 | f
echo "== 8K replicate N=$N"; timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 48 --warmup 5 --replicate --no-verify 2>&1 | f
echo "== 16K peer N=$N"; timeout 600 $TR --master-port 29515 bench.py --gpus $N --steps 48 --warmup 5 --width 15360 --height 8640 --no-verify 2>&1 | f
echo "== config 5 N=$N"; timeout 600 $TR --master-port 29516 bench.py --gpus $N --config 5 --steps 48 --warmup 5 2>&1 | f
} > gpurun_out/suite_n$N.log 2>&1
cat gpurun_out/suite_n$N.log
