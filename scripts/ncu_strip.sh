#!/bin/bash
# full ncu captures of the strip kernel with the register caps named by $SHAPES ("minb ..."), config $CFG
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for mb in ${SHAPES:-3}; do
  TAA_STRIP_MINB=$mb timeout 900 ncu --set full --clock-control none --import-source on -k regex:taa_resolve_strip -s ${SKIP:-4} -c 1 -f \
    -o gpurun_out/prof_strip_c${CFG:-2}_m${mb} python bench.py --kernel-only --config ${CFG:-2} --steps 8 --warmup 4 > gpurun_out/ncu_strip.log 2>&1
  tail -2 gpurun_out/ncu_strip.log
done
