#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest chain"; timeout 1700 python -m pytest tests/test_chain_gpu.py -m gpu -q --timeout=1600 --tb=short 2>&1 | grep -E "^(FAILED|ERROR|E  )|passed|failed" | cut -c1-300 | head -30
echo "== fused"; timeout 300 python scripts/debug/fused_time.py 2>&1 | tail -1
echo "== unfused"; TAA_FUSED_CHAIN=0 timeout 300 python scripts/debug/fused_time.py 2>&1 | tail -1
echo "== bench"; timeout 900 python bench.py --steps 100 --warmup 5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('ms_per_step','varying_motion','config3_full_chain','config3_resolve_only','fused_resolve_cas')})"
} > gpurun_out/r2j.log 2>&1
cat gpurun_out/r2j.log
