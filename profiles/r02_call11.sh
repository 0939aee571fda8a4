#!/usr/bin/env bash
# round-2 GPU call 11 (4 GPUs): grouped view sharding (2 groups x 2 ranks) on real NCCL ranks, the two NCCL tests, and
# bench.py at N=4 with its extra workloads; every step under its own short timeout
set -u
mkdir -p gpurun_out/r02
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
nvidia-smi -L | wc -l
SCENES=2 timeout 240 $TR --master-port 29551 tests/run_viewshard.py > gpurun_out/r02/viewshard_4gpu.log 2>&1; echo "viewshard rc=$?"
grep -E "VIEWSHARD|capture" gpurun_out/r02/viewshard_4gpu.log || tail -n 20 gpurun_out/r02/viewshard_4gpu.log
timeout 500 python -m pytest tests/test_viewshard_gpu.py tests/test_temporal_gpu.py -m gpu -q > gpurun_out/r02/pytest_2gpu.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r02/pytest_2gpu.log
timeout 420 $TR --master-port 29554 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02/bench_n4.json 2> gpurun_out/r02/bench_n4.err; echo "bench n4 rc=$?"
cat gpurun_out/r02/bench_n4.json; tail -n 5 gpurun_out/r02/bench_n4.err
