#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== memcheck smoke"; timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "Host Frame" | head -80
for t in "test_single_frame[config2]" "test_single_frame[config3]" "test_each_switch_of_the_family[mRejectOutside=1]" "test_each_switch_of_the_family[mDepthCulling=1]" "test_each_switch_of_the_family[mDynamicAntiGhosting=1]" "test_tiny_and_ragged_sizes[size0]" "test_tiny_and_ragged_sizes[size5]" "test_reset_history" "test_smoothly_varying_motion_runs_the_general_strip_path[config2]" "test_uniform_motion_tiles[pan0]" "test_row_bands_equal_whole_frame[2]"; do
echo "== $t"; timeout 300 python -m pytest "tests/test_tuned_gpu.py::$t" -m gpu -q --timeout=200 --tb=short 2>&1 | grep -E "Error|error|assert|passed|failed" | head -8
done
} > gpurun_out/diag.log 2>&1
tail -150 gpurun_out/diag.log
