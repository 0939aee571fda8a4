#!/usr/bin/env bash
# round-2 GPU call 20 (1 GPU): neighboring_attn_type concat / self, per-view prompts (use_aug_text) -- full GPU suite
set -u
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02/pytest_gpu_3.log 2>&1; echo "pytest rc=$?"
tail -n 25 gpurun_out/r02/pytest_gpu_3.log
