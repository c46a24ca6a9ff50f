#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest parity"; timeout 1500 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout=600 --tb=short -k defaults 2>&1 | grep -E "^(FAILED|ERROR|E  )|passed|failed" | cut -c1-260 | head
timeout 300 python scripts/debug/cfg1_time.py 2>&1 | tail -1
for mb in 4 5 6; do echo "== fused minb $mb"; TAA_STREAM_EPI_MINB=$mb timeout 300 python scripts/debug/fused_time.py 2>&1 | tail -1; done
echo "== fused minb 4 tail 0"; TAA_STREAM_TAIL=0 TAA_STREAM_EPI_MINB=4 timeout 300 python scripts/debug/fused_time.py 2>&1 | tail -1
echo "== fused R=26"; TAA_STREAM_R=26 timeout 300 python scripts/debug/fused_time.py 2>&1 | tail -1
echo "== fused R=20"; TAA_STREAM_R=20 timeout 300 python scripts/debug/fused_time.py 2>&1 | tail -1
} > gpurun_out/r2x.log 2>&1
cat gpurun_out/r2x.log
