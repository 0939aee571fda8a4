set -u
mkdir -p gpurun_out/r02
for t in old new old new; do
  if [ $t = new ]; then lib=dualdiff_b200/libdualdiff_sm100.so; else lib=profiles/ab/lib_$t.so; fi
  timeout 600 python profiles/bench_with_lib.py $lib --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-library-baseline 2> gpurun_out/r02/bench_call38_$t.err | tee gpurun_out/r02/bench_call38_$t.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$t', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['frac'], {k:v['ms'] for k,v in d['kernel_breakdown'].items()})"
done
