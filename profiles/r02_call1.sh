#!/usr/bin/env bash
# round-2 GPU call 1: suite sanity on this round's box, then validate + time the round-1 experimental switches
set -u
mkdir -p gpurun_out/r02
python -m pytest tests -m gpu -x -q > gpurun_out/r02/pytest_gpu_0.log 2>&1; echo "pytest rc=$?"
DD_EXPERIMENTAL=1 python -m pytest tests/test_experimental.py -m gpu -q > gpurun_out/r02/pytest_experimental.log 2>&1; echo "experimental rc=$?"
for cfg in "base" "DD_CONV_IN_PATCH=1" "DD_SMALL_CONV_IM2COL=28" "DD_CONV_IN_PATCH=1 DD_SMALL_CONV_IM2COL=28"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  if [ "$cfg" = base ]; then envs=""; else envs="$cfg"; fi
  env $envs DD_BENCH_SHAPES=gpurun_out/r02/shapes_$tag.txt python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02/bench_$tag.json 2> gpurun_out/r02/bench_$tag.err
  echo "bench $tag rc=$?"; head -c 300 gpurun_out/r02/bench_$tag.json; echo
done
