#!/bin/bash
# quick GPU session: parity tests of the default path, smoke, kernel-only lines (pan and varying motion), the fused chain
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
b() { timeout 300 python bench.py --kernel-only --steps 100 --warmup 6 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['frac'], d['motion'])"; }
{
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== pytest tuned+chain+streams"; timeout 1500 python -m pytest tests/test_tuned_gpu.py tests/test_chain_gpu.py tests/test_streams.py -m gpu -q --timeout=900 --tb=short 2>&1 | grep -E "^(FAILED|ERROR|E  )|passed|failed" | cut -c1-260 | head -30
echo "-- config 2, pan"; b
echo "-- config 2, varying motion"; b --motion varying
echo "-- config 3"; b --config 3
echo "-- fused resolve + CAS"; timeout 300 python scripts/debug/fused_time.py 2>&1 | tail -1
} > gpurun_out/quick.log 2>&1
cat gpurun_out/quick.log
