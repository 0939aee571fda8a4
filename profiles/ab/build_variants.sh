#!/usr/bin/env bash
# Same-box A/B builds of the CUDA library (after dualdiff_b200/csrc/build.sh):
#   profiles/ab/lib_old.so      dd_gemm.cu / dd_attention.cu / dd_common.cuh of git revision $OLD_REV (default HEAD), the rest as built
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
for f in dd_gemm.cu dd_attention.cu dd_common.cuh dd_api_internal.h; do git show $OLD_REV:dualdiff_b200/csrc/$f > build/old/$f; done
sed -i 's#"../../include/dualdiff_b200.h"#"../../../../include/dualdiff_b200.h"#' build/old/dd_api_internal.h
( $NVCC $FLAGS -c build/old/dd_gemm.cu -o build/dd_gemm_old.o &
  $NVCC $FLAGS -c build/old/dd_attention.cu -o build/old/dd_attention.o &
  wait
  $NVCC -shared -o ../../profiles/ab/lib_old.so build/dd_gemm_old.o build/old/dd_attention.o $(ls build/dd_*.o | grep -v "dd_gemm\|dd_attention") \
    -lcudart_static -lpthread -ldl -lrt && echo "built profiles/ab/lib_old.so" ) &
for v in "$@"; do
  case $v in
    noareuse) build_one noareuse -DDD_GEMM_AREUSE=0 & ;;
    nomerged) build_one nomerged -DDD_CONV_MERGED=0 & ;;
    probe*)   build_one $v -DDD_GEMM_AREUSE=0 -DDD_PROBE=${v#probe} & ;;
  esac
done
wait
