#!/bin/bash
# Round-2 A/B pass (one gpurun call): L2 prefetch in align_fast_kernel_pe, host chunk size of the e2e path.
set -u
O=gpurun_out
mkdir -p $O
for pf in 0 1; do
  GIRAFFE_B200_FAST_PREFETCH=$pf python scripts/profile_pe.py 1000000 4 > $O/exp_prefetch_$pf.log 2>&1
  echo "prefetch=$pf: $(grep -h 'kernel ms' $O/exp_prefetch_$pf.log | tail -1 | cut -c1-400)"
done
GIRAFFE_B200_FAST_PREFETCH=1 python -m pytest tests/test_map_paired_parity.py tests/test_vcf_graphs.py -m gpu -q > $O/exp_prefetch_tests.log 2>&1; echo "prefetch tests rc=$? $(tail -1 $O/exp_prefetch_tests.log)"
for chunk in 524288 1048576 2097152; do
  for pf in 0 1; do
    GIRAFFE_B200_FAST_PREFETCH=$pf GIRAFFE_B200_MAP_CHUNK=$chunk python bench.py --steps 3 --warmup 2 --cpu-seconds 1 --no-secondary > $O/exp_chunk_${chunk}_$pf.json 2> /dev/null
    python - <<PY
import json
d=json.loads(open("$O/exp_chunk_${chunk}_$pf.json").read().strip().splitlines()[-1])
print("chunk=$chunk prefetch=$pf device %.2f M  e2e %.2f M  parity %s" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["config"]["parity_vs_cpu_sample"]))
PY
  done
done
