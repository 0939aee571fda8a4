#!/usr/bin/env bash
# round-2 GPU call 14 (1 GPU): ones-column softmax denominator -- full GPU suite, level-0 A/B (self: v_ones vs own row sum in one
# process), bench, racecheck re-run
set -u
mkdir -p gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02/pytest_gpu_2.log 2>&1; echo "pytest rc=$?"
tail -n 6 gpurun_out/r02/pytest_gpu_2.log
timeout 300 python profiles/attn_ones_ab.py > gpurun_out/r02/attn_ones_ab.txt 2>&1; echo "ab rc=$?"; cat gpurun_out/r02/attn_ones_ab.txt
DD_BENCH_SHAPES=gpurun_out/r02/shapes_call14.txt timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-library-baseline > gpurun_out/r02/bench_call14.json 2> gpurun_out/r02/bench_call14.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02/bench_call14.json'))
print(d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['kernel_breakdown'].items()})
PY
grep attn gpurun_out/r02/shapes_call14.txt | head -6
timeout 400 compute-sanitizer --tool racecheck --print-limit 20 --log-file gpurun_out/r02/sanitizer_racecheck2.log python profiles/memcheck_step.py 1 8 12 > gpurun_out/r02/sanitizer_racecheck2.out 2>&1; echo "racecheck rc=$?"
tail -n 4 gpurun_out/r02/sanitizer_racecheck2.log; tail -n 6 gpurun_out/r02/sanitizer_racecheck2.out | cut -c1-300
