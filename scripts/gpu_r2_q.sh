#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
b() { timeout 300 python bench.py --kernel-only --steps 100 --warmup 6 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['frac'], d['motion'])"; }
{
echo "== pytest tuned"; timeout 1500 python -m pytest tests/test_tuned_gpu.py -m gpu -q --timeout=600 --tb=short 2>&1 | grep -E "^(FAILED|ERROR|E  )|passed|failed" | cut -c1-260 | head -30
echo "-- pan default"; b
echo "-- varying default"; b --motion varying
} > gpurun_out/r2q.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:taa_resolve -s 8 -c 1 -f -o gpurun_out/r2q_prof_var python bench.py --kernel-only --motion varying --steps 8 --warmup 4 > gpurun_out/r2q_ncu_var.log 2>&1
cat gpurun_out/r2q.log
