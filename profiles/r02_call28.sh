#!/usr/bin/env bash
# round-2 GPU call 28 (1 GPU): 3x3 conv with one ring slot per (kernel row, 64 channels) -- one barrier wait + one commit per
# twelve UMMAs on tiles up to 192 wide: parity, per-shape timing and issuer timeline against the separate-ring build, whole-step A/B
set -u
mkdir -p gpurun_out/r02
timeout 420 python -m pytest tests/test_gemm_gpu.py -x -q > gpurun_out/r02/pytest_call28.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r02/pytest_call28.log
for t in profiles/ab/lib_nomerged.so dualdiff_b200/libdualdiff_sm100.so; do timeout 300 python profiles/conv_probe.py $t; done 2>&1 | tee gpurun_out/r02/conv_probe.txt
timeout 200 python profiles/gemm_trace.py conv > gpurun_out/r02/gemm_trace_conv3.txt 2>&1; grep -E "^===|^launch|issuer, cycles|tile committed|period" gpurun_out/r02/gemm_trace_conv3.txt | cut -c1-160 | head -30
for t in old new nomerged new; do
  if [ $t = new ]; then lib=dualdiff_b200/libdualdiff_sm100.so; else lib=profiles/ab/lib_$t.so; fi
  DD_BENCH_SHAPES=gpurun_out/r02/shapes_call28_$t.txt timeout 600 python profiles/bench_with_lib.py $lib --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-library-baseline 2> gpurun_out/r02/bench_call28_$t.err | tee gpurun_out/r02/bench_call28_$t.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$t', d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['kernel_breakdown'].items()})"
done
