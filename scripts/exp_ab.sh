#!/bin/bash
# A/B of library builds on one box: build/exp/lib*.so, same bench + secondary runs for each
for L in "$@"; do
  export GIRAFFE_B200_LIB=$PWD/build/exp/lib$L.so
  python bench.py --steps 2 --warmup 1 --cpu-seconds 1 --no-secondary --reads 4000000 > gpurun_out/exp_$L.json 2> gpurun_out/exp_$L.err
  python -c "
import json; d=json.load(open('gpurun_out/exp_$L.json')); k=d['roofline']['kernel_ms_last_chunk']; print('lib$L', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), {a:round(b,2) for a,b in k.items() if b>0.2}, d['config']['parity_vs_cpu_sample'])"
  for c in config5 config4 config2se; do
    python tests/tools/run_config.py $c 200000 2>&1 | tail -1 | python -c "
import sys,ast
l=sys.stdin.read(); i=l.index('['); k=ast.literal_eval(l[i:].strip()); d={}
for a,b in k: d[a]=d.get(a,0)+b
print('  ', l[:i].strip(), {a:round(b,2) for a,b in d.items() if b>0.25})"
  done
done
unset GIRAFFE_B200_LIB
