#!/usr/bin/env bash
# round-2 GPU call 26 (1 GPU): TMA epilogue with in-place residual tiles loaded a chunk ahead, chunks beyond N skipped, the odd
# chunk alternating between the warp halves, carried tile coordinates: parity, per-shape timing against the previous
# commit's kernel, CTA timeline, whole-step A/B
set -u
mkdir -p gpurun_out/r02
timeout 420 python -m pytest tests/test_gemm_gpu.py -x -q > gpurun_out/r02/pytest_call26.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r02/pytest_call26.log
for t in profiles/ab/lib_old.so profiles/ab/lib_noareuse.so dualdiff_b200/libdualdiff_sm100.so; do timeout 300 python profiles/gemm_probe.py $t check; done 2>&1 | tee gpurun_out/r02/gemm_probe3.txt
timeout 200 python profiles/gemm_trace.py > gpurun_out/r02/gemm_trace2.txt 2>&1; echo "trace rc=$?"
for t in old new noareuse; do
  if [ $t = new ]; then lib=dualdiff_b200/libdualdiff_sm100.so; else lib=profiles/ab/lib_$t.so; fi
  DD_BENCH_SHAPES=gpurun_out/r02/shapes_call26_$t.txt timeout 600 python profiles/bench_with_lib.py $lib --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-library-baseline 2> gpurun_out/r02/bench_call26_$t.err | tee gpurun_out/r02/bench_call26_$t.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$t', d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['kernel_breakdown'].items()})"
done
