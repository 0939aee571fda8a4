#!/usr/bin/env bash
# round-2 GPU call 2: new level-0 attention kernel -- parity tests, then A/B timing of the three variants
set -u
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention or head_dim_40 or box_features or conv_in" > gpurun_out/r02/pytest_attn.log 2>&1; echo "pytest rc=$?"
tail -n 15 gpurun_out/r02/pytest_attn.log
timeout 300 python profiles/attn_one.py > gpurun_out/r02/attn_ab.txt 2>&1; echo "attn_one rc=$?"
cat gpurun_out/r02/attn_ab.txt
