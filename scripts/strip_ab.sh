#!/bin/bash
# A/B of the strip kernel (TAA_STRIP_MINB = CTAs/SM the registers are capped for, TAA_STRIP_UNROLL = strip loop unrolling) against the
# 32x32-tile kernel, device-timed only; then the parity tests.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for cfg in ${CFGS:-2 3}; do
  [ -z "$NOTILE" ] && { echo "cfg $cfg tile"; TAA_TUNED_VARIANT=tile python bench.py --kernel-only --config $cfg --steps 200 --warmup 10 2>&1 | tail -1; }
  for mb in ${MINBS:-2 3}; do for un in ${UNROLLS:-4}; do
    echo "cfg $cfg strip minb $mb unroll $un"; TAA_STRIP_MINB=$mb TAA_STRIP_UNROLL=$un python bench.py --kernel-only --config $cfg --steps 200 --warmup 10 2>&1 | tail -1
  done; done
done
[ -z "$NOTESTS" ] && { echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -15; }
} 2>&1 | tee gpurun_out/strip_ab.log
if [ -n "$NCU_SHAPES" ]; then SHAPES="$NCU_SHAPES" bash scripts/ncu_strip.sh; fi
