#!/usr/bin/env bash
# round-2 GPU call 21 (1 GPU): the driver's sequence -- GPU suite, smoke(), reference arm, default bench (with extras, library and CPU
# baselines) -- plus a fresh ncu capture of the level-0 attention kernel
set -u
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02/pytest_gpu_4.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/r02/pytest_gpu_4.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 3
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02/bench_reference.json 2> gpurun_out/r02/bench_reference.err; echo "reference rc=$?"; cat gpurun_out/r02/bench_reference.json
SECONDS=0
DD_BENCH_SHAPES=gpurun_out/r02/shapes_final.txt timeout 1200 python bench.py > gpurun_out/r02/bench_final.json 2> gpurun_out/r02/bench_final.err; echo "bench rc=$? wall ${SECONDS}s"
cat gpurun_out/r02/bench_final.json; tail -n 3 gpurun_out/r02/bench_final.err
N_IMG=24 VARIANT=0 ONLY=self V_ONES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_pp -s 3 -c 1 -o gpurun_out/r02/attn_pp_ones python profiles/attn_one.py > gpurun_out/r02/ncu_attn_pp_ones.log 2>&1; echo "ncu rc=$?"
