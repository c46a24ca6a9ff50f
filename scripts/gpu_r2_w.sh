#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest parity (defaults kernel)"; timeout 1500 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout=600 --tb=short 2>&1 | grep -E "^(FAILED|ERROR|E  )|passed|failed" | cut -c1-260 | head -30
echo "== bench"; timeout 900 python bench.py --steps 100 --warmup 5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('ms_per_step','varying_motion','config1_defaults','config3_full_chain','fused_resolve_cas')})"
echo "== defaults off"; TAA_DEFAULTS_KERNEL=0 timeout 900 python bench.py --steps 40 --warmup 5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config1_defaults'])"
} > gpurun_out/r2w.log 2>&1
cat gpurun_out/r2w.log
