#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest tuned"; timeout 1500 python -m pytest tests/test_tuned_gpu.py -m gpu -q --timeout=600 --tb=short 2>&1 | grep -E "^(FAILED|ERROR|E  )|passed|failed" | cut -c1-260 | head -60
echo "== cfg2 pan auto"; timeout 300 python bench.py --kernel-only --steps 100 --warmup 5 2>&1 | tail -1
for r in 14 19 26 30; do echo "== cfg2 pan R=$r"; TAA_STREAM_R=$r timeout 300 python bench.py --kernel-only --steps 100 --warmup 5 2>&1 | tail -1; done
echo "== cfg3 pan"; timeout 300 python bench.py --kernel-only --config 3 --steps 100 --warmup 5 2>&1 | tail -1
echo "== cfg2 varying"; timeout 300 python bench.py --kernel-only --motion varying --steps 100 --warmup 5 2>&1 | tail -1
} > gpurun_out/r2e.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:taa_resolve -s 8 -c 1 -f -o gpurun_out/r2e_prof python bench.py --kernel-only --steps 8 --warmup 4 > gpurun_out/r2e_ncu_full.log 2>&1
cat gpurun_out/r2e.log
