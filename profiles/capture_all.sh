set -x
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r01_bench_reference.json 2>gpurun_out/ref.err
DD_BENCH_SHAPES=gpurun_out/r01_step_shapes.txt python bench.py --steps 10 --warmup 3 > gpurun_out/r01_bench.json 2>gpurun_out/bench.err
EAGER=1 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r01_launches.csv python profiles/step_once.py > gpurun_out/step_once.log 2>&1
LEVEL=0 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/r01_conv_l0 python profiles/conv_one.py > /dev/null 2>&1
LEVEL=2 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/r01_conv_l2 python profiles/conv_one.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/r01_gemm_n320_k320 python profiles/gemm_one.py > /dev/null 2>&1
GN=2560 GRES=0 ncu --set full --clock-control none -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/r01_gemm_n2560_k320 python profiles/gemm_one.py > /dev/null 2>&1
N_IMG=96 ncu --set full --clock-control none --import-source on -k regex:attn_v2 -s 2 -c 1 -o gpurun_out/r01_attn_l0 python profiles/attn_one.py > /dev/null 2>&1
ncu --set full --clock-control none -k regex:gn_ -s 20 -c 2 -o gpurun_out/r01_gn_c320 python profiles/gn_one.py > /dev/null 2>&1
python profiles/pipeline_e2e.py 8 25 > gpurun_out/r01_pipeline_e2e.json 2>gpurun_out/pipe.err
EXTRAS=1 compute-sanitizer --tool memcheck --print-limit 30 --log-file gpurun_out/memcheck_full.log python profiles/memcheck_step.py 2 28 50 2>gpurun_out/memcheck_full.err
python profiles/gap_table.py gpurun_out/r01_step_shapes.txt > gpurun_out/r01_gap_table.txt
# experimental variants: parity, then A/B against the default build on the same box
DD_EXPERIMENTAL=1 python -m pytest tests/test_experimental.py -m gpu -q > gpurun_out/experimental_tests.log 2>&1
DD_CONV_IN_PATCH=1 python bench.py --steps 10 --warmup 3 > gpurun_out/ab_conv_in_patch.json 2>/dev/null
DD_SMALL_CONV_IM2COL=28 python bench.py --steps 10 --warmup 3 > gpurun_out/ab_small_conv_im2col.json 2>/dev/null
ls -la gpurun_out | tail -20

# ---- round 2 (final state): the driver's sequence + evidence behind profiles/r02_*; see r02_call39.sh / r02_call40.sh for the runs as made
bash profiles/r02_call39.sh          # GPU suite, smoke, reference arm, default bench, ncu launch list + ncu --set full (L0 conv, K = 320 GEMM), pipeline timing
bash profiles/r02_call40.sh          # compute-sanitizer memcheck (full latent) + synccheck
# same-box A/B and timelines (CPU box first: bash dualdiff_b200/csrc/build.sh && bash profiles/ab/build_variants.sh [noareuse nomerged], then
#  nvcc -DDD_GEMM_TRACE / -DDD_ATTN_TRACE builds linked to profiles/ab/lib_trace.so / lib_attn_trace.so as in profiles/README.md):
#   python profiles/gemm_probe.py <lib> check ; python profiles/conv_probe.py <lib> ; python profiles/attn_probe.py <lib> ; python profiles/gn_probe.py <lib>
#   python profiles/gemm_trace.py [conv] ; python profiles/attn_trace.py 2 text ; bash profiles/ab/ab_bench.sh ; bash profiles/ab/ab_streams.sh
