#!/bin/bash
# N GPUs: sharded == whole frame, then the sharded bench with peer stores and with the NCCL exchange
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
{
echo "== sharded_check N=$N"; timeout 600 $TR --master-port 29511 scripts/sharded_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -8
echo "== bench peer N=$N"; timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 96 --warmup 5 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -2
echo "== bench nccl N=$N"; TAA_SHARDED_EXCHANGE=nccl timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 96 --warmup 5 --no-verify 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -2
} > gpurun_out/multi_n$N.log 2>&1
cat gpurun_out/multi_n$N.log
