#!/bin/bash
# A/B of the tuned-kernel variants (TAA_TUNED_VARIANT bit 0: L1 warm-up, bit 1: 4 CTAs/SM register cap), device-timed only.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in ${CFGS:-2 3}; do for v in ${VARIANTS:-0 1 2 3}; do
  TAA_TUNED_VARIANT=$v python bench.py --kernel-only --config $cfg --steps 200 --warmup 10 2>&1 | tail -1
done; done | tee gpurun_out/variants.log
