#!/usr/bin/env bash
# Same-box A/B builds of the CUDA library (after dualdiff_b200/csrc/build.sh):
#   profiles/ab/lib_old.so      every source of git revision $OLD_REV (default HEAD)
#   profiles/ab/lib_<tag>.so    the working-tree sources with -D overrides for dd_gemm.cu (DD_GEMM_AREUSE selects the tile
#                               schedule; DD_PROBE the differential-timing probes, whose results are wrong by construction)
# Run one of them with  python profiles/bench_with_lib.py profiles/ab/lib_<tag>.so ...  or  profiles/gemm_probe.py <lib>.
set -euo pipefail
cd "$(dirname "$0")/../../dualdiff_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
OLD_REV=${OLD_REV:-HEAD}
others=$(ls build/dd_*.o | grep -v "dd_gemm\|dd_attention_")
link() { $NVCC -shared -o ../../profiles/ab/lib_$1.so $2 $others -lcudart_static -lpthread -ldl -lrt; echo "built profiles/ab/lib_$1.so"; }
build_one() {  # tag, defines...
  local tag=$1; shift
  $NVCC $FLAGS "$@" -c dd_gemm.cu -o build/dd_gemm_$tag.o
  link $tag build/dd_gemm_$tag.o
}
mkdir -p build/old
rm -f build/old/*
for f in $(git ls-tree --name-only $OLD_REV ./ | grep -E "\.(cu|cuh|h)$"); do git show $OLD_REV:./$f > build/old/$f; done
sed -i 's#"../../include/dualdiff_b200.h"#"../../../../include/dualdiff_b200.h"#' build/old/*.cu build/old/*.h build/old/*.cuh
( for f in build/old/dd_*.cu; do $NVCC $FLAGS -c $f -o ${f%.cu}.o & done
  wait
  $NVCC -shared -o ../../profiles/ab/lib_old.so build/old/dd_*.o -lcudart_static -lpthread -ldl -lrt &&
    python3 -c "import ctypes; ctypes.CDLL('../../profiles/ab/lib_old.so').dd_launch_count" && echo "built profiles/ab/lib_old.so" ) &
for v in "$@"; do
  case $v in
    noareuse) build_one noareuse -DDD_GEMM_AREUSE=0 & ;;
    nomerged) build_one nomerged -DDD_CONV_MERGED=0 & ;;
    probe*)   build_one $v -DDD_GEMM_AREUSE=0 -DDD_PROBE=${v#probe} & ;;
  esac
done
wait
