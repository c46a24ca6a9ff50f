#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest chain+fxaa"; timeout 1500 python -m pytest tests/test_chain_gpu.py tests/test_fxaa_gpu.py -m gpu -q --timeout=900 --tb=short 2>&1 | grep -E "^(FAILED|ERROR|E  )|passed|failed" | cut -c1-260 | head -20
timeout 300 python scripts/debug/post_time.py 2>&1 | tail -3
} > gpurun_out/post.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"cas_kernel|sharpen_kernel|post_process_kernel|fxaa|sharpen_rows" -s 18 -c 6 -f -o gpurun_out/prof_post python scripts/debug/post_time.py > gpurun_out/ncu_post.log 2>&1
cat gpurun_out/post.log; tail -2 gpurun_out/ncu_post.log
