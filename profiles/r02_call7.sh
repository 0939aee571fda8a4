#!/usr/bin/env bash
set -u
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02/pytest_gpu_1.log 2>&1; echo "pytest rc=$?"
tail -n 5 gpurun_out/r02/pytest_gpu_1.log
grep -E "after step|final latents|scene . of|guided eps at|raw eps sample|bg branch without|12 down residuals|branch . mid" gpurun_out/r02/pytest_gpu_1.log
DD_BENCH_SHAPES=gpurun_out/r02/shapes_attn.txt timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02/bench_attn.json 2> gpurun_out/r02/bench_attn.err; echo "bench rc=$?"
head -c 400 gpurun_out/r02/bench_attn.json; echo
