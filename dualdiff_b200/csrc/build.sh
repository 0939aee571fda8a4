#!/usr/bin/env bash
# Build libdualdiff_sm100.so in-tree (nvcc cross-compiles for sm_100a without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
OUT=../libdualdiff_sm100.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
mkdir -p build
objs=()
for f in dd_*.cu; do
  o=build/${f%.cu}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ dd_common.cuh -nt "$o" ] || [ dd_api_internal.h -nt "$o" ] || [ ../../include/dualdiff_b200.h -nt "$o" ]; then
    echo "nvcc $f"
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c "$f" -o "$o" &
  fi
  objs+=("$o")
done
wait
$NVCC -shared -o $OUT "${objs[@]}" -lcudart_static -lpthread -ldl -lrt
echo "built $OUT"
