#!/usr/bin/env bash
set -u
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention or head_dim_40" > gpurun_out/r02/pytest_attn.log 2>&1; echo "pytest rc=$?"
tail -n 5 gpurun_out/r02/pytest_attn.log
timeout 300 python profiles/attn_one.py > gpurun_out/r02/attn_ab3.txt 2>&1; echo "attn_one rc=$?"
cat gpurun_out/r02/attn_ab3.txt
N_IMG=24 VARIANT=2 ONLY=self timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_pp -s 3 -c 1 -o gpurun_out/r02/attn_pp_v3 python profiles/attn_one.py > gpurun_out/r02/ncu_attn_pp.log 2>&1; echo "ncu rc=$?"
