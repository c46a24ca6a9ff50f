#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for nw in 1 2; do cp build/ab/lib_nw$nw.so taa_star_b200/libtaa_b200.so; echo "== NWARP $nw"
MOVER=1 timeout 300 python scripts/debug/stream_trace.py 2>&1 | tail -9 | cut -c1-700
MOVER=0 timeout 300 python scripts/debug/stream_trace.py 2>&1 | tail -9 | head -1
for r in 30; do MOVER=1 TAA_STREAM_R=$r timeout 300 python scripts/debug/stream_trace.py 2>&1 | tail -9 | head -1; done
timeout 300 python bench.py --kernel-only --steps 100 --warmup 6 2>&1 | tail -1
timeout 300 python bench.py --kernel-only --steps 100 --warmup 6 --motion varying 2>&1 | tail -1
done
} > gpurun_out/r2l.log 2>&1
cat gpurun_out/r2l.log
