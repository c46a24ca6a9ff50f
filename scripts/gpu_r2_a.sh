#!/bin/bash
# Round 2, first GPU session of the streaming kernel: smoke, default-path parity tests, kernel-only timings (A/B against the strip kernel),
# memcheck of one small frame, launch list and one full ncu capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== probe"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest tuned"; timeout 1200 python -m pytest tests/test_tuned_gpu.py -m gpu -q --timeout=600 --tb=line 2>&1 | tail -40
for cfg in 2 3; do for mo in pan varying; do
echo "== stream cfg$cfg $mo"; timeout 300 python bench.py --kernel-only --config $cfg --motion $mo --steps 100 --warmup 5 2>&1 | tail -1
echo "== strip  cfg$cfg $mo"; TAA_TUNED_VARIANT=strip timeout 300 python bench.py --kernel-only --config $cfg --motion $mo --steps 100 --warmup 5 2>&1 | tail -1
done; done
for r in 14 19 24 30; do echo "== stream cfg2 pan R=$r"; TAA_STREAM_R=$r timeout 300 python bench.py --kernel-only --steps 100 --warmup 5 2>&1 | tail -1; done
echo "== memcheck"; timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15
} > gpurun_out/r2a.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:taa_ -c 30 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --kernel-only --steps 10 --warmup 3 > gpurun_out/r2a_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:taa_resolve -s 8 -c 1 -f -o gpurun_out/r2a_prof python bench.py --kernel-only --steps 8 --warmup 4 > gpurun_out/r2a_ncu_full.log 2>&1
tail -80 gpurun_out/r2a.log
