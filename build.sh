#!/bin/bash
# Builds libtaa_b200.so (sm_100a only) in-tree. No GPU needed: nvcc cross-compiles.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
SRC=taa_star_b200/csrc
OUT=taa_star_b200/libtaa_b200.so
ARCH="-gencode arch=compute_100a,code=sm_100a"
# --fmad=false: the EXACT kernels must not contract a*b+c (see taa_device.cuh); FAST kernels opt in with explicit fmaf.
FLAGS="-O3 -std=c++17 -lineinfo --fmad=false -Xcompiler -fPIC,-fvisibility=hidden -ccbin /usr/bin/g++ ${EXTRA_NVCC_FLAGS:-}"
mkdir -p build
objs=()
pids=()
for f in $SRC/*.cu; do
  o=build/$(basename "${f%.cu}").o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find $SRC include -newer "$o" \( -name '*.h' -o -name '*.cuh' \) -print -quit)" ]; then
    echo "nvcc $f"
    $NVCC $ARCH $FLAGS -c "$f" -o "$o" &
    pids+=($!)
  fi
  objs+=("$o")
done
for p in "${pids[@]:-}"; do [ -z "$p" ] || wait "$p" || { echo "build.sh: a compilation failed" >&2; exit 1; }; done
$NVCC $ARCH -shared -o $OUT "${objs[@]}" -Xcompiler -fPIC -ccbin /usr/bin/g++
echo "built $OUT"
