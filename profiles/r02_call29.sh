#!/usr/bin/env bash
# round-2 GPU call 29 (1 GPU): register epilogue without the per-row 64-bit division (image index carried with the row), chunk
# halves alternating per tile: parity, conv timing against the previous commit, timeline, whole-step A/B
set -u
mkdir -p gpurun_out/r02
timeout 420 python -m pytest tests/test_gemm_gpu.py -x -q > gpurun_out/r02/pytest_call29.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r02/pytest_call29.log
for t in profiles/ab/lib_old.so dualdiff_b200/libdualdiff_sm100.so; do timeout 300 python profiles/conv_probe.py $t; done 2>&1 | tee gpurun_out/r02/conv_probe2.txt
timeout 200 python profiles/gemm_trace.py conv > gpurun_out/r02/gemm_trace_conv4.txt 2>&1; grep -E "^===|^launch|issuer, cycles|tile committed|period|transposed|stored" gpurun_out/r02/gemm_trace_conv4.txt | cut -c1-160 | head -40
for t in old new old new; do
  if [ $t = new ]; then lib=dualdiff_b200/libdualdiff_sm100.so; else lib=profiles/ab/lib_$t.so; fi
  DD_BENCH_SHAPES=gpurun_out/r02/shapes_call29_$t.txt timeout 600 python profiles/bench_with_lib.py $lib --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-library-baseline 2> gpurun_out/r02/bench_call29_$t.err | tee gpurun_out/r02/bench_call29_$t.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$t', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], {k:v['ms'] for k,v in d['kernel_breakdown'].items()})"
done
