#!/bin/bash
# GPU parity tests only (optionally a -k filter as $1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=600 ${1:+-k "$1"} 2>&1 | tail -40 | tee gpurun_out/tests.log
