#!/usr/bin/env bash
# round-2 GPU call 40 (1 GPU): compute-sanitizer over one sampler step on the final kernels (memcheck at the full 28x50 latent with
# the side streams, synccheck on the small ragged step) -- the GEMM epilogues and the conv ring were rewritten this session
set -u
mkdir -p gpurun_out/r02
timeout 240 compute-sanitizer --tool memcheck --print-limit 20 --log-file gpurun_out/r02/sanitizer2_memcheck_full.log python profiles/memcheck_step.py 2 28 50 > gpurun_out/r02/sanitizer2_memcheck_full.out 2>&1; echo "memcheck full rc=$?"
tail -n 2 gpurun_out/r02/sanitizer2_memcheck_full.log; tail -n 2 gpurun_out/r02/sanitizer2_memcheck_full.out | cut -c1-200
timeout 200 compute-sanitizer --tool synccheck --print-limit 20 --log-file gpurun_out/r02/sanitizer2_synccheck.log python profiles/memcheck_step.py 1 8 12 > gpurun_out/r02/sanitizer2_synccheck.out 2>&1; echo "synccheck rc=$?"
tail -n 2 gpurun_out/r02/sanitizer2_synccheck.log; tail -n 2 gpurun_out/r02/sanitizer2_synccheck.out | cut -c1-200
