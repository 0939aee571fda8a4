#!/usr/bin/env bash
# round-2 GPU call 12 (2 GPUs): token-sharded temporal attention (all-to-all) -- kernel test, 2-rank parity, frameshard bench
set -u
mkdir -p gpurun_out/r02
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_temporal_gpu.py -m gpu -q -x > gpurun_out/r02/pytest_temporal.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/r02/pytest_temporal.log
FRAMES=4 VIDEO_STEP=1 VIDEO_UNET=1 timeout 300 $TR --master-port 29561 tests/run_frameshard.py > gpurun_out/r02/frameshard_2gpu_a2a.log 2>&1; echo "frameshard rc=$?"
grep -E "FRAMESHARD|VIDEOSTEP|VIDEOUNET|capture" gpurun_out/r02/frameshard_2gpu_a2a.log || tail -n 30 gpurun_out/r02/frameshard_2gpu_a2a.log
timeout 300 $TR --master-port 29562 bench.py --gpus 2 --steps 5 --warmup 3 --workload frameshard > gpurun_out/r02/bench_frameshard_n2.json 2> gpurun_out/r02/bench_frameshard_n2.err; echo "bench rc=$?"
cat gpurun_out/r02/bench_frameshard_n2.json; tail -n 3 gpurun_out/r02/bench_frameshard_n2.err
