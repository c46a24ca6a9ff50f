#!/bin/bash
# quick GPU session on 2 GPUs: the default sharded bench line (with the whole-frame check) + the GPU parity tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
{
echo "== bench default N=2"; timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 96 --warmup 5 2>&1 | grep "^{\|Error\|whole-frame" | cut -c1-2500
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short 2>&1 | grep -E "^(FAILED|ERROR|E  )|passed|failed" | cut -c1-260 | head -20
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
} > gpurun_out/quick.log 2>&1
cat gpurun_out/quick.log
