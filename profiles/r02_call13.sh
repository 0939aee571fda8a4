#!/usr/bin/env bash
# round-2 GPU call 13 (1 GPU): GroupNorm row-split clusters -- parity, A/B table, ncu; sanitizer passes (memcheck, racecheck,
# synccheck) over one small sampler step; bench
set -u
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "groupnorm" > gpurun_out/r02/pytest_gn2.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/r02/pytest_gn2.log
timeout 300 python profiles/gn_one.py > gpurun_out/r02/gn_ab2.txt 2>&1; echo "gn_one rc=$?"; cat gpurun_out/r02/gn_ab2.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-library-baseline > gpurun_out/r02/bench_call13.json 2> gpurun_out/r02/bench_call13.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02/bench_call13.json'))
print(d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['kernel_breakdown'].items()})
PY
GN_C=320 GN_H=28 GN_W=50 TWO_PASS=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gn_slab -s 3 -c 1 -o gpurun_out/r02/gn_cluster_c320 python profiles/gn_one.py > gpurun_out/r02/ncu_gn2.log 2>&1; echo "ncu rc=$?"
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 --log-file gpurun_out/r02/sanitizer_$tool.log python profiles/memcheck_step.py 1 8 12 > gpurun_out/r02/sanitizer_$tool.out 2>&1; echo "$tool rc=$?"
  tail -n 3 gpurun_out/r02/sanitizer_$tool.log
done
