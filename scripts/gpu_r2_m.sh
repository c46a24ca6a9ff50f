#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest tuned+chain"; timeout 1500 python -m pytest tests/test_tuned_gpu.py tests/test_chain_gpu.py -m gpu -q --timeout=600 --tb=short 2>&1 | grep -E "^(FAILED|ERROR|E  )|passed|failed" | cut -c1-260 | head -60
for hnt in 1 0; do for m in 1 0; do echo "== hints $hnt mover $m"; TAA_STREAM_HINTS=$hnt MOVER=$m timeout 300 python scripts/debug/stream_trace.py 2>&1 | tail -9 | cut -c1-900; done; done
echo "== cfg2 varying"; timeout 300 python bench.py --kernel-only --motion varying --steps 100 --warmup 5 2>&1 | tail -1
echo "== fused"; timeout 300 python scripts/debug/fused_time.py 2>&1 | tail -1
} > gpurun_out/r2m.log 2>&1
cat gpurun_out/r2m.log
