#!/bin/bash
cd "$(dirname "$0")/.."
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
f() { grep -v "OMP_NUM\|^W1018\|^\*\*\*" | grep "AssertionError\|^{\|whole-frame check\|sharded_check\|Error" | head -6 | cut -c1-3000; }
{
echo "== sharded_check N=$N"; timeout 600 $TR --master-port 29511 scripts/sharded_check.py 2>&1 | f
echo "== 8K peer N=$N"; timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 96 --warmup 5 2>&1 | f
} > gpurun_out/final_n8.log 2>&1
cat gpurun_out/final_n8.log | cut -c1-900
