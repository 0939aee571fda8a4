#!/usr/bin/env bash
# round-2 GPU call 9 (2 GPUs): the two sharded configurations on real NCCL ranks -- parity scripts, the two NCCL tests,
# then bench.py at N=2 (main line + extra_workloads) and at N=1 (sub_metrics, library_gpu_baseline)
set -u
mkdir -p gpurun_out/r02
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi -L
timeout 600 $TR --master-port 29541 tests/run_viewshard.py > gpurun_out/r02/viewshard_2gpu.log 2>&1; echo "viewshard rc=$?"
grep -E "VIEWSHARD|capture" gpurun_out/r02/viewshard_2gpu.log
LATENT=56x100 SCENES=2 timeout 600 $TR --master-port 29542 tests/run_viewshard.py > gpurun_out/r02/viewshard_2gpu_hd.log 2>&1; echo "viewshard hd rc=$?"
grep -E "VIEWSHARD|capture" gpurun_out/r02/viewshard_2gpu_hd.log
FRAMES=4 VIDEO_STEP=1 timeout 600 $TR --master-port 29543 tests/run_frameshard.py > gpurun_out/r02/frameshard_2gpu.log 2>&1; echo "frameshard rc=$?"
grep -E "FRAMESHARD|VIDEOSTEP|capture" gpurun_out/r02/frameshard_2gpu.log
timeout 900 python -m pytest tests/test_viewshard_gpu.py tests/test_temporal_gpu.py -m gpu -q > gpurun_out/r02/pytest_2gpu.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r02/pytest_2gpu.log
timeout 900 $TR --master-port 29544 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02/bench_n2.json 2> gpurun_out/r02/bench_n2.err; echo "bench n2 rc=$?"
cat gpurun_out/r02/bench_n2.json; tail -n 5 gpurun_out/r02/bench_n2.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02/bench_n1.json 2> gpurun_out/r02/bench_n1.err; echo "bench n1 rc=$?"
cat gpurun_out/r02/bench_n1.json; tail -n 5 gpurun_out/r02/bench_n1.err
