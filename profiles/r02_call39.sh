#!/usr/bin/env bash
# round-2 GPU call 39 (1 GPU): the driver's sequence on the final kernels -- GPU suite, smoke(), reference arm, default bench (with
# extras, library and CPU baselines) -- then the ncu launch list of one step, ncu --set full of the L0 conv and the K = 320 GEMM,
# pipeline timing
set -u
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02/pytest_gpu_5.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/r02/pytest_gpu_5.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 3
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02/bench_reference2.json 2> gpurun_out/r02/bench_reference2.err; echo "reference rc=$?"; cat gpurun_out/r02/bench_reference2.json
SECONDS=0
DD_BENCH_SHAPES=gpurun_out/r02/shapes_final2.txt timeout 1200 python bench.py > gpurun_out/r02/bench_final2.json 2> gpurun_out/r02/bench_final2.err; echo "bench rc=$? wall ${SECONDS}s"
cat gpurun_out/r02/bench_final2.json; tail -n 3 gpurun_out/r02/bench_final2.err
EAGER=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02/r02_launches2.csv python profiles/step_once.py > gpurun_out/r02/step_once2.log 2>&1; echo "launch list rc=$?"
LEVEL=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/r02/r02_conv_l0 python profiles/conv_one.py > gpurun_out/r02/ncu_conv_l0.log 2>&1; echo "ncu conv rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/r02/r02_gemm_n320_k320 python profiles/gemm_one.py > gpurun_out/r02/ncu_gemm_k320.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 python profiles/pipeline_e2e.py 8 25 > gpurun_out/r02/r02_pipeline_e2e2.json 2> gpurun_out/r02/pipe2.err; echo "pipeline rc=$?"; cat gpurun_out/r02/r02_pipeline_e2e2.json
