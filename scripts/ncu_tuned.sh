#!/bin/bash
# one full ncu capture of the tuned resolve kernel (+ its fix-up) out of a short kernel-only bench run
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:taa_resolve -s ${SKIP:-8} -c ${COUNT:-2} -f -o gpurun_out/prof_tuned \
  python bench.py --kernel-only --config ${CFG:-2} --steps 8 --warmup 4 > gpurun_out/ncu_tuned.log 2>&1
tail -3 gpurun_out/ncu_tuned.log
