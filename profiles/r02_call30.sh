#!/usr/bin/env bash
# round-2 GPU call 30 (1 GPU): register epilogue without the per-row 64-bit division (vector row carried with the output row),
# chunk halves alternating per tile: GEMM + step parity, whole-step A/B against the previous commit
set -u
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_step_gpu.py tests/test_benchmarked_gpu.py -x -q > gpurun_out/r02/pytest_call30.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r02/pytest_call30.log
for t in old new old new; do
  if [ $t = new ]; then lib=dualdiff_b200/libdualdiff_sm100.so; else lib=profiles/ab/lib_$t.so; fi
  DD_BENCH_SHAPES=gpurun_out/r02/shapes_call30_$t.txt timeout 600 python profiles/bench_with_lib.py $lib --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-library-baseline 2> gpurun_out/r02/bench_call30_$t.err | tee gpurun_out/r02/bench_call30_$t.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$t', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['frac'], {k:v['ms'] for k,v in d['kernel_breakdown'].items()})"
done
