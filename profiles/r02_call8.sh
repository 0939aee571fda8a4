#!/usr/bin/env bash
set -u
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_boundary_gpu.py tests/test_vae.py -x -q -k "groupnorm or boundary or module or box_tokens or topology or vae" > gpurun_out/r02/pytest_gn.log 2>&1; echo "pytest rc=$?"
tail -n 12 gpurun_out/r02/pytest_gn.log
timeout 300 python profiles/gn_one.py > gpurun_out/r02/gn_ab.txt 2>&1; echo "gn_one rc=$?"
cat gpurun_out/r02/gn_ab.txt
