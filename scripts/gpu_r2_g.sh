#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest all gpu"; timeout 1700 python -m pytest tests -m gpu -q --timeout=1600 --tb=short 2>&1 | grep -E "^(FAILED|ERROR|E  )|passed|failed" | cut -c1-300 | head -40
echo "== cfg3 stream"; TAA_STREAM_REJ=1 timeout 300 python bench.py --kernel-only --config 3 --steps 100 --warmup 5 2>&1 | tail -1
} > gpurun_out/r2g.log 2>&1
TAA_STREAM_REJ=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:taa_resolve_stream -s 8 -c 1 -f -o gpurun_out/r2g_prof_c3 python bench.py --kernel-only --config 3 --steps 8 --warmup 4 > gpurun_out/r2g_ncu.log 2>&1
cat gpurun_out/r2g.log
