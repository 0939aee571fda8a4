# same-box A/B: the two condition branches on side streams (default) against one stream
set -u
mkdir -p gpurun_out/r02
for t in par ser par ser; do
  if [ $t = ser ]; then flag="--serial-branches"; else flag=""; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-library-baseline $flag 2> gpurun_out/r02/bench_streams_$t.err | tee gpurun_out/r02/bench_streams_$t.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$t', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['config']['branch_streams'])"
done
