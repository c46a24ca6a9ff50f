#!/bin/bash
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
f() { grep -v "OMP_NUM\|^W1018\|^\*\*\*" | grep "AssertionError\|^{\|debug\|whole-frame check" | head -30 | cut -c1-900; }
{
echo "== peer verify"; timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 96 --warmup 5 2>&1 | f
} > gpurun_out/r2t_n$N.log 2>&1
cat gpurun_out/r2t_n$N.log
