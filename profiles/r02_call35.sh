#!/usr/bin/env bash
# round-2 GPU call 35 (1 GPU): tcgen05 fence / accumulator release before the final global stores (attention epilogues, GEMM
# register epilogue): parity, per-shape timing, whole-step A/B against the previous commit
set -u
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py tests/test_step_gpu.py -x -q -k "not groupnorm and not layernorm" > gpurun_out/r02/pytest_call35.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r02/pytest_call35.log
for t in profiles/ab/lib_old.so dualdiff_b200/libdualdiff_sm100.so; do timeout 300 python profiles/attn_probe.py $t; timeout 300 python profiles/conv_probe.py $t; done 2>&1 | tee gpurun_out/r02/probe_call35.txt
for t in old new old new; do
  if [ $t = new ]; then lib=dualdiff_b200/libdualdiff_sm100.so; else lib=profiles/ab/lib_$t.so; fi
  DD_BENCH_SHAPES=gpurun_out/r02/shapes_call35_$t.txt timeout 600 python profiles/bench_with_lib.py $lib --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-library-baseline 2> gpurun_out/r02/bench_call35_$t.err | tee gpurun_out/r02/bench_call35_$t.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$t', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['frac'], {k:v['ms'] for k,v in d['kernel_breakdown'].items()})"
done
