#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
b() { timeout 300 python bench.py --kernel-only --steps 100 --warmup 6 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['frac'], d['motion'])"; }
{
echo "== pytest tuned"; timeout 1500 python -m pytest tests/test_tuned_gpu.py -m gpu -q --timeout=600 --tb=short 2>&1 | grep -E "^(FAILED|ERROR|E  )|passed|failed" | cut -c1-260 | head -30
for h in 0 1; do for t in 0 15 25 35 50; do for rs in 14 8; do echo "-- hints $h tail $t rs $rs"; TAA_STREAM_HINTS=$h TAA_STREAM_TAIL=$t TAA_STREAM_RS=$rs b; done; done; done
for r in 20 30; do for t in 25 40; do echo "-- hints 1 R $r tail $t rs 10"; TAA_STREAM_R=$r TAA_STREAM_TAIL=$t TAA_STREAM_RS=10 b; done; done
for pf in 0 33 34 17 18 49 50; do echo "-- varying genpf $pf"; TAA_STREAM_GENPF=$pf b --motion varying; done
for pf in 33 34; do echo "-- pan hints 1 tail 25 genpf $pf"; TAA_STREAM_TAIL=25 TAA_STREAM_GENPF=$pf b; done
} > gpurun_out/r2n.log 2>&1
cat gpurun_out/r2n.log
