#!/usr/bin/env bash
# round-2 GPU call 15 (1 GPU): GroupNorm single statistics pass (pivoted moments, packed math) -- parity, timing table, ncu;
# attention with the ones column: FMA-pipe share on / off; ncu of the LayerNorm kernel; bench
set -u
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_vae.py -x -q -k "groupnorm or vae" > gpurun_out/r02/pytest_gn3.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/r02/pytest_gn3.log
timeout 300 python profiles/gn_one.py > gpurun_out/r02/gn_ab3.txt 2>&1; echo "gn_one rc=$?"; grep "one-pass" gpurun_out/r02/gn_ab3.txt
timeout 300 python profiles/attn_ones_ab.py > gpurun_out/r02/attn_ones_ab2.txt 2>&1; echo "ab rc=$?"; cat gpurun_out/r02/attn_ones_ab2.txt
DD_BENCH_SHAPES=gpurun_out/r02/shapes_call15.txt timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-library-baseline > gpurun_out/r02/bench_call15.json 2> gpurun_out/r02/bench_call15.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02/bench_call15.json'))
print(d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['kernel_breakdown'].items()})
PY
GN_C=320 GN_H=28 GN_W=50 TWO_PASS=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn_slab -s 3 -c 1 -o gpurun_out/r02/gn_slab2_c320 python profiles/gn_one.py > gpurun_out/r02/ncu_gn3.log 2>&1; echo "ncu gn rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:layernorm -s 2 -c 1 -o gpurun_out/r02/layernorm_c320 python profiles/ln_one.py > gpurun_out/r02/ncu_ln.log 2>&1; echo "ncu ln rc=$?"
timeout 120 python profiles/ln_one.py > gpurun_out/r02/ln_one.txt 2>&1; cat gpurun_out/r02/ln_one.txt
