#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
b() { timeout 300 python bench.py --kernel-only --steps 100 --warmup 6 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['frac'], d['motion'])"; }
{
echo "-- pan default"; b
echo "-- varying default"; b --motion varying
echo "-- varying hints 0"; TAA_STREAM_HINTS=0 b --motion varying
echo "-- varying hints 0 R 30"; TAA_STREAM_HINTS=0 TAA_STREAM_R=30 b --motion varying
} > gpurun_out/r2o.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:taa_resolve -s 8 -c 1 -f -o gpurun_out/r2o_prof_pan python bench.py --kernel-only --steps 8 --warmup 4 > gpurun_out/r2o_ncu_pan.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:taa_resolve -s 8 -c 1 -f -o gpurun_out/r2o_prof_var python bench.py --kernel-only --motion varying --steps 8 --warmup 4 > gpurun_out/r2o_ncu_var.log 2>&1
cat gpurun_out/r2o.log
