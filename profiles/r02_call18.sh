#!/usr/bin/env bash
# round-2 GPU call 18 (1 GPU): GEMM issuer with barrier probes one stage ahead -- parity, then same-box A/B of the whole step
# (before / after / before / after)
set -u
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py tests/test_step_gpu.py -x -q > gpurun_out/r02/pytest_gemm.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r02/pytest_gemm.log
for rep in 1 2; do
  for which in before after; do
    if [ $which = before ]; then
      DD_BENCH_SHAPES=gpurun_out/r02/shapes_gemm_${which}.txt timeout 600 python profiles/bench_with_lib.py profiles/ab/lib_gemm_before.so --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-library-baseline > gpurun_out/r02/bench_gemm_${which}_$rep.json 2>/dev/null
    else
      DD_BENCH_SHAPES=gpurun_out/r02/shapes_gemm_${which}.txt timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-library-baseline > gpurun_out/r02/bench_gemm_${which}_$rep.json 2>/dev/null
    fi
    python - <<PY
import json
d=json.load(open('gpurun_out/r02/bench_gemm_${which}_$rep.json'))
print('$which', $rep, d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['kernel_breakdown'].items()})
PY
  done
done
for k in M134400_N320_K2880 M134400_N320_K320 M33600_N640_K640 M134400_N1152_K320 M8736_N1280_K11520 M2688_N1280_K11520; do grep -h $k gpurun_out/r02/shapes_gemm_before.txt gpurun_out/r02/shapes_gemm_after.txt; done
