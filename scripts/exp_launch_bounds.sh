for cfg in "4 0" "3 0" "2 0" "4 12" "4 16"; do
  set -- $cfg
  GIRAFFE_B200_EXTEND_MINB=$1 GIRAFFE_B200_FAST_MINB=$2 python bench.py --steps 2 --warmup 1 --cpu-seconds 1 --no-secondary --reads 4000000 > gpurun_out/exp.json 2> gpurun_out/exp.err
  python -c "
import json; d=json.load(open('gpurun_out/exp.json')); k=d['roofline']['kernel_ms_last_chunk']; print('EXTEND_MINB=$1 FAST_MINB=$2', round(d['value']/1e6,2), 'extend', k['extend_kernel'], 'fast', k['align_fast_kernel_pe'], d['config']['parity_vs_cpu_sample'])"
done
timeout 600 compute-sanitizer --tool racecheck --print-limit 12 python -m pytest tests/test_map_paired_parity.py tests/test_xdrop_golden.py -m gpu -q -k "rescue or vectors or repeats" > gpurun_out/r02_sanitizer_racecheck.log 2>&1
grep -A12 "Warning\|Error" gpurun_out/r02_sanitizer_racecheck.log | head -60 | cut -c1-220
