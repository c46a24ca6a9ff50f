#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest tuned+chain"; timeout 1500 python -m pytest tests/test_tuned_gpu.py tests/test_chain_gpu.py -m gpu -q --timeout=600 --tb=short 2>&1 | grep -E "^(FAILED|ERROR|E  )|passed|failed" | cut -c1-260 | head -60
echo "== cfg2 pan auto"; timeout 300 python bench.py --kernel-only --steps 100 --warmup 5 2>&1 | tail -1
for r in 26 39 52 78 90; do echo "== cfg2 pan R=$r"; TAA_STREAM_R=$r timeout 300 python bench.py --kernel-only --steps 100 --warmup 5 2>&1 | tail -1; done
echo "== cfg2 varying"; timeout 300 python bench.py --kernel-only --motion varying --steps 100 --warmup 5 2>&1 | tail -1
echo "== cfg2 varying R=26"; TAA_STREAM_R=26 timeout 300 python bench.py --kernel-only --motion varying --steps 100 --warmup 5 2>&1 | tail -1
echo "== fused"; timeout 300 python scripts/debug/fused_time.py 2>&1 | tail -1
echo "== fused R=26"; TAA_STREAM_R=26 timeout 300 python scripts/debug/fused_time.py 2>&1 | tail -1
echo "== unfused"; TAA_FUSED_CHAIN=0 timeout 300 python scripts/debug/fused_time.py 2>&1 | tail -1
} > gpurun_out/r2k.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:taa_resolve -s 8 -c 1 -f -o gpurun_out/r2k_prof python bench.py --kernel-only --steps 8 --warmup 4 > gpurun_out/r2k_ncu_full.log 2>&1
cat gpurun_out/r2k.log
