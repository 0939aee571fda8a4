#!/usr/bin/env bash
# round-2 GPU call 10 (1 GPU): persistent level 1-3 attention + reordered barrier waits + SFA+ -- parity tests, A/B timing,
# TMEM / MUFU microbenchmarks, bench (3 streams vs 1), ncu of the level-0 attention kernel
set -u
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_boundary_gpu.py -x -q > gpurun_out/r02/pytest_call10.log 2>&1; echo "pytest rc=$?"
tail -n 6 gpurun_out/r02/pytest_call10.log
timeout 300 python profiles/attn_one.py > gpurun_out/r02/attn_ab4.txt 2>&1; echo "attn_one rc=$?"; cat gpurun_out/r02/attn_ab4.txt
timeout 300 python profiles/attn_levels.py > gpurun_out/r02/attn_levels.txt 2>&1; echo "attn_levels rc=$?"; cat gpurun_out/r02/attn_levels.txt
timeout 120 profiles/micro/tmem_bench.bin > gpurun_out/r02/tmem_bench.txt 2>&1; echo "tmem rc=$?"; cat gpurun_out/r02/tmem_bench.txt
timeout 120 profiles/micro/mufu_bench2.bin > gpurun_out/r02/mufu_bench2.txt 2>&1; echo "mufu rc=$?"; cat gpurun_out/r02/mufu_bench2.txt
DD_BENCH_SHAPES=gpurun_out/r02/shapes_call10.txt timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02/bench_call10.json 2> gpurun_out/r02/bench_call10.err; echo "bench rc=$?"
cat gpurun_out/r02/bench_call10.json; tail -n 5 gpurun_out/r02/bench_call10.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-library-baseline --serial-branches > gpurun_out/r02/bench_call10_serial.json 2> gpurun_out/r02/bench_call10_serial.err; echo "bench serial rc=$?"
head -c 600 gpurun_out/r02/bench_call10_serial.json; echo
N_IMG=24 VARIANT=0 ONLY=self timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_pp -s 3 -c 1 -o gpurun_out/r02/attn_pp_r02 python profiles/attn_one.py > gpurun_out/r02/ncu_attn_pp_r02.log 2>&1; echo "ncu rc=$?"
