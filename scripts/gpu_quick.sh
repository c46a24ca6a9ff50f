#!/bin/bash
# quick GPU session: parity tests of the default path, kernel-only A/B lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
b() { timeout 300 python bench.py --kernel-only --steps 100 --warmup 6 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['frac'], d['motion'])"; }
{
echo "== pytest tuned+chain"; timeout 1500 python -m pytest tests/test_tuned_gpu.py tests/test_chain_gpu.py -m gpu -q --timeout=900 --tb=short 2>&1 | grep -E "^(FAILED|ERROR|E  )|passed|failed" | cut -c1-260 | head -20
echo "-- pan default"; b
echo "-- pan minb 7 (144 regs)"; TAA_STREAM_MINB=7 b
echo "-- varying default"; b --motion varying
echo "-- varying minb 7"; TAA_STREAM_MINB=7 b --motion varying
echo "== pytest tuned under minb 7"; TAA_STREAM_MINB=7 timeout 1500 python -m pytest tests/test_tuned_gpu.py -m gpu -q --timeout=900 --tb=short -k "single_frame or 64_frame or full_size or without_a_mask or bands" 2>&1 | grep -E "^(FAILED|ERROR|E  )|passed|failed" | cut -c1-260 | head
} > gpurun_out/quick.log 2>&1
cat gpurun_out/quick.log
