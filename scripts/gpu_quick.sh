#!/bin/bash
# Short GPU session: parity tests, smoke, kernel-only bench lines for configs 2 and 3.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== probe"; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv; nproc
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -15
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== bench cfg2 kernel-only"; timeout 600 python bench.py --kernel-only --steps 100 --warmup 5 2>&1 | tail -2
echo "== bench cfg3 kernel-only"; timeout 600 python bench.py --kernel-only --config 3 --steps 100 --warmup 5 2>&1 | tail -2
} > gpurun_out/quick.log 2>&1
tail -40 gpurun_out/quick.log
