#!/bin/bash
# One GPU session: probe, parity tests, smoke, bench, ncu launch list + full capture of the resolve kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== probe"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv; nproc
ls /usr/share/vulkan/icd.d 2>/dev/null; which glslangValidator glslc vulkaninfo 2>/dev/null; echo "vulkan probe done"
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -40
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== bench"; timeout 600 python bench.py --steps 100 --warmup 5 2>&1 | tail -5
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -3
} > gpurun_out/round.log 2>&1
if [ "${NCU:-1}" = "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:taa_ -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 10 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:taa_resolve -s 8 -c 2 -f -o gpurun_out/prof_resolve python bench.py --kernel-only --steps 8 --warmup 4 > gpurun_out/ncu_full.log 2>&1
fi
tail -60 gpurun_out/round.log
if [ "${NCU:-1}" = "1" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:taa_resolve -s 8 -c 2 -f -o gpurun_out/prof_resolve_c3 python bench.py --kernel-only --config 3 --steps 8 --warmup 4 > gpurun_out/ncu_full_c3.log 2>&1
fi
