#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for m in 1 0; do for r in 26 20 15; do MOVER=$m TAA_STREAM_R=$r timeout 300 python scripts/debug/stream_trace.py 2>&1 | tail -9; done; done
} > gpurun_out/r2l.log 2>&1
cat gpurun_out/r2l.log
