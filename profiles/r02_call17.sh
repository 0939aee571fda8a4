#!/usr/bin/env bash
# round-2 GPU call 17 (1 GPU): LayerNorm at 4 CTAs/SM, GroupNorm back on 256 threads; ncu launch list of one step; pipeline e2e
set -u
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "groupnorm or layernorm" > gpurun_out/r02/pytest_ln2.log 2>&1; echo "pytest rc=$?"
tail -n 2 gpurun_out/r02/pytest_ln2.log
timeout 300 python profiles/gn_one.py 2>&1 | grep "one-pass" | head -3
timeout 120 python profiles/ln_one.py > gpurun_out/r02/ln_one3.txt 2>&1; cat gpurun_out/r02/ln_one3.txt
DD_BENCH_SHAPES=gpurun_out/r02/shapes_call17.txt timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-library-baseline > gpurun_out/r02/bench_call17.json 2> gpurun_out/r02/bench_call17.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02/bench_call17.json'))
print(d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['kernel_breakdown'].items()})
PY
EAGER=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02/r02_launches.csv python profiles/step_once.py > gpurun_out/r02/step_once.log 2>&1; echo "launch list rc=$?"
timeout 600 python profiles/pipeline_e2e.py 8 25 > gpurun_out/r02/r02_pipeline_e2e.json 2> gpurun_out/r02/pipe.err; echo "pipeline rc=$?"; cat gpurun_out/r02/r02_pipeline_e2e.json
