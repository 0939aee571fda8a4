#!/usr/bin/env bash
# round-2 GPU call 22 (N GPUs): the driver's multi-GPU launch of the default bench (main line + view- / frame-sharded extras)
set -u
N=${N:-2}
mkdir -p gpurun_out/r02
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L | wc -l
SECONDS=0
timeout 600 $TR --master-port 29571 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02/bench_final_n$N.json 2> gpurun_out/r02/bench_final_n$N.err; echo "bench n$N rc=$? wall ${SECONDS}s"
python - <<PY
import json
d=json.load(open('gpurun_out/r02/bench_final_n$N.json'))
print('main', d['value'], d['ms_per_step'], d['config']['parallelism'])
for k,v in d.get('extra_workloads',{}).items():
    print(k, v if 'value' not in v else (v['value'], v['ms_per_step'], v['config']['cuda_graph'], v['config']['parallelism'][:80]))
PY
tail -n 4 gpurun_out/r02/bench_final_n$N.err
if [ "$N" = "2" ]; then
  SCENES=2 LATENT=28x50 timeout 300 $TR --master-port 29572 tests/run_viewshard.py 2>&1 | grep -E "VIEWSHARD|capture|Error" | head -5
fi
