#!/usr/bin/env bash
# round-2 GPU call 16 (1 GPU): packed LayerNorm, 512-thread GroupNorm slabs -- parity, timing tables, bench, ncu of LayerNorm
set -u
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_vae.py tests/test_clip.py -x -q -k "groupnorm or vae or layernorm or clip or attention or head_dim" > gpurun_out/r02/pytest_ln.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/r02/pytest_ln.log
timeout 300 python profiles/gn_one.py > gpurun_out/r02/gn_ab4.txt 2>&1; echo "gn_one rc=$?"; grep "one-pass" gpurun_out/r02/gn_ab4.txt
timeout 120 python profiles/ln_one.py > gpurun_out/r02/ln_one2.txt 2>&1; cat gpurun_out/r02/ln_one2.txt
DD_BENCH_SHAPES=gpurun_out/r02/shapes_call16.txt timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-library-baseline > gpurun_out/r02/bench_call16.json 2> gpurun_out/r02/bench_call16.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02/bench_call16.json'))
print(d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['kernel_breakdown'].items()})
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:layernorm -s 2 -c 1 -o gpurun_out/r02/layernorm_packed_c320 python profiles/ln_one.py > gpurun_out/r02/ncu_ln2.log 2>&1; echo "ncu ln rc=$?"
