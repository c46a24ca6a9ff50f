#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for r in 1 3; do
for d in 0 64 0 64; do echo "dbg $d"; BANDS=8 RANK_ID=$r PEER=1 TAA_PEER_DEBUG=$d timeout 200 python scripts/debug/band_time.py 2>&1 | tail -1 | cut -c1-80; done
done
} > gpurun_out/quick.log 2>&1
cat gpurun_out/quick.log
