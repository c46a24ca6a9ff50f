#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 200 python scripts/debug/band_time.py 2>&1 | tail -1
for t in 0 10 30; do TAA_STREAM_TAIL=$t timeout 200 python scripts/debug/band_time.py 2>&1 | tail -1; done
TAA_STREAM_HINTS=0 timeout 200 python scripts/debug/band_time.py 2>&1 | tail -1
for r in 12 13 14 16 18 20 22 24 27 30; do TAA_STREAM_TAIL=0 TAA_STREAM_R=$r timeout 200 python scripts/debug/band_time.py 2>&1 | tail -1; done
for r in 13 14 20 27; do TAA_STREAM_TAIL=25 TAA_STREAM_R=$r timeout 200 python scripts/debug/band_time.py 2>&1 | tail -1; done
BANDS=4 timeout 200 python scripts/debug/band_time.py 2>&1 | tail -1
BANDS=4 TAA_STREAM_TAIL=0 timeout 200 python scripts/debug/band_time.py 2>&1 | tail -1
} > gpurun_out/r2v.log 2>&1
cat gpurun_out/r2v.log
