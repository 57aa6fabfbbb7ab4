#!/bin/bash
# extend_kernel launch-bounds sweep on one library build: exp_minb.sh <lib letter> <minb>...
L=$1; shift
export GIRAFFE_B200_LIB=$PWD/build/exp/lib$L.so
for M in "$@"; do
  GIRAFFE_B200_EXTEND_MINB=$M python bench.py --steps 2 --warmup 1 --cpu-seconds 1 --no-secondary --reads 4000000 > gpurun_out/exp_${L}_$M.json 2> gpurun_out/exp_${L}_$M.err
  python -c "
import json; d=json.load(open('gpurun_out/exp_${L}_$M.json')); k=d['roofline']['kernel_ms_last_chunk']; print('lib$L EXTEND_MINB=$M', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), {a:round(b,2) for a,b in k.items() if b>0.2}, d['config']['parity_vs_cpu_sample'])"
done
