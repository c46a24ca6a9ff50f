#!/bin/bash
# Round 2: smoke, default-path parity tests, kernel-only timings of the streaming kernel against the strip kernel, one full ncu capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest tuned"; timeout 1500 python -m pytest tests/test_tuned_gpu.py -m gpu -q --timeout=600 --tb=short 2>&1 | grep -E "^(FAILED|ERROR|E  )|passed|failed" | cut -c1-260 | head -80
for cfg in 2 3; do for mo in pan varying; do
echo "== stream cfg$cfg $mo"; timeout 300 python bench.py --kernel-only --config $cfg --motion $mo --steps 100 --warmup 5 2>&1 | tail -1
done; done
for r in 14 19 24; do echo "== stream cfg2 pan R=$r"; TAA_STREAM_R=$r timeout 300 python bench.py --kernel-only --steps 100 --warmup 5 2>&1 | tail -1; done
echo "== memcheck"; timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "Host Frame" | tail -25
} > gpurun_out/r2b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:taa_resolve -s 8 -c 1 -f -o gpurun_out/r2b_prof python bench.py --kernel-only --steps 8 --warmup 4 > gpurun_out/r2b_ncu_full.log 2>&1
cat gpurun_out/r2b.log
