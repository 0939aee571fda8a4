#!/usr/bin/env bash
# round-2 GPU call 23 (1 GPU): GEMM tile schedule candidates (A reuse over contiguous tile ranges, L2 prefetch of the next
# tile's A boxes), differential probes of the TMA-epilogue GEMMs, query-tile L2 prefetch in the attention kernels
set -u
mkdir -p gpurun_out/r02
timeout 420 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py -x -q -k "not groupnorm and not layernorm" > gpurun_out/r02/pytest_call23.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r02/pytest_call23.log
for t in base areuse pfa; do timeout 300 python profiles/gemm_probe.py profiles/ab/lib_$t.so check; done 2>&1 | tee gpurun_out/r02/gemm_probe.txt
timeout 300 python profiles/gemm_probe.py dualdiff_b200/libdualdiff_sm100.so check 2>&1 | tee -a gpurun_out/r02/gemm_probe.txt
for t in probe1 probe2 probe3; do timeout 300 python profiles/gemm_probe.py profiles/ab/lib_$t.so; done 2>&1 | tee -a gpurun_out/r02/gemm_probe.txt
for t in ab/lib_attn_nopfq.so ../dualdiff_b200/libdualdiff_sm100.so; do timeout 300 python profiles/attn_probe.py profiles/$t; done 2>&1 | tee gpurun_out/r02/attn_probe.txt
for t in old new; do
  if [ $t = old ]; then lib=profiles/ab/lib_old.so; else lib=dualdiff_b200/libdualdiff_sm100.so; fi
  DD_BENCH_SHAPES=gpurun_out/r02/shapes_call23_$t.txt timeout 600 python profiles/bench_with_lib.py $lib --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-library-baseline 2> gpurun_out/r02/bench_call23_$t.err | tee gpurun_out/r02/bench_call23_$t.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$t', d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['kernel_breakdown'].items()})"
done
