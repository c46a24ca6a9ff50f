#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest chain+fxaa"; timeout 1700 python -m pytest tests/test_chain_gpu.py tests/test_fxaa_gpu.py -m gpu -q --timeout=1600 --tb=short 2>&1 | grep -E "^(FAILED|ERROR|E  )|passed|failed" | cut -c1-300 | head -40
echo "== pytest rest"; timeout 1700 python -m pytest tests -m gpu -q --timeout=1600 --tb=short --deselect tests/test_chain_gpu.py --deselect tests/test_fxaa_gpu.py 2>&1 | grep -E "^(FAILED|ERROR|E  )|passed|failed" | cut -c1-300 | head -20
} > gpurun_out/r2h.log 2>&1
cat gpurun_out/r2h.log
