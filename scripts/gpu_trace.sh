#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
MOVER=0 timeout 300 python scripts/debug/stream_trace.py 2>&1 | tail -9 | cut -c1-900
MOVER=0 TAA_STREAM_PERSIST=0 timeout 300 python scripts/debug/stream_trace.py 2>&1 | tail -9 | head -4| cut -c1-900
} > gpurun_out/trace.log 2>&1
cat gpurun_out/trace.log
