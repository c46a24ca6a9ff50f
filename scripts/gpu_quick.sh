#!/bin/bash
# quick GPU session: all parity tests, the bench line, a few A/B lines given as "ENV=.. ENV=.. -- bench args" in $AB (one per line)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --tb=short 2>&1 | grep -E "^(FAILED|ERROR|E  )|passed|failed" | cut -c1-260 | head -30
echo "== bench"; timeout 900 python bench.py --steps 100 --warmup 5 2>/dev/null | tail -1 > gpurun_out/quick_bench_line.json; python -c "
import json; d=json.load(open('gpurun_out/quick_bench_line.json')); print({k:d[k] for k in ('ms_per_step','value','e2e','gpu_launches')}); print({k:(d[k]['ms_per_step'],d[k]['frac']) for k in ('varying_motion','config1_defaults','config3_full_chain','config3_resolve_only','fused_resolve_cas')})"
for mb in 4 5 6; do echo "== cfg3 stream-rej minb $mb"; TAA_STREAM_REJ=1 TAA_STREAM_MINB=$mb timeout 300 python bench.py --kernel-only --config 3 --steps 100 --warmup 5 2>&1 | tail -1 | cut -c1-200; done
echo "== cfg3 strip"; timeout 300 python bench.py --kernel-only --config 3 --steps 100 --warmup 5 2>&1 | tail -1 | cut -c1-200
} > gpurun_out/quick.log 2>&1
cat gpurun_out/quick.log
