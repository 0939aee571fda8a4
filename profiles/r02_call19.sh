#!/usr/bin/env bash
# round-2 GPU call 19 (1 GPU): level-1 attention with 32-key tiles on 3 CTAs per SM (variant 2) against the 64-key / 2-CTA form
set -u
mkdir -p gpurun_out/r02
timeout 300 python profiles/attn_levels.py > gpurun_out/r02/attn_levels2.txt 2>&1; echo "attn_levels rc=$?"; grep "d=80" gpurun_out/r02/attn_levels2.txt
