#!/bin/bash
# Round-2 measurement pass on one B200 (run under gpurun from the repo root):
#   gpurun --timeout 2400 -- 'bash scripts/measure_r02.sh'
# Everything lands in gpurun_out/; scripts/profile_summary.py turns the ncu outputs into profiles/*.{json,md} afterwards.
set -u
T0=$SECONDS; lap() { echo "[t+$((SECONDS-T0))s] $*"; }
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/r02_gputest.log 2>&1; lap "gpu tests rc=$?"; tail -2 $O/r02_gputest.log
python bench.py > $O/r02_bench_head.json 2> $O/r02_bench_head.err; lap "bench rc=$?"
K='seed_kernel|extend_|align_|tail_plan|tail_decide|xdrop_tile|prep_pairs|compact_gather|rebase_offsets|advance_run|DeviceScan'
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 800 --csv --log-file $O/r02_launches.csv \
    python bench.py --steps 1 --warmup 1 --cpu-seconds 1 --no-secondary > $O/r02_bench_under_ncu.log 2>&1; lap "launch list rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"seed_kernel_pe|extend_|align_fast|align_kernel|tail_plan|tail_decide|xdrop_tile" -c 40 \
    -o $O/r02_full_pe python scripts/profile_pe.py 1000000 2 > $O/r02_full_pe.log 2>&1; lap "full pe rc=$?"
# gpurun brings back at most 64 MiB: keep the raw metric page (and the hottest source lines), not the report
ncu -i $O/r02_full_pe.ncu-rep --page raw --csv > $O/r02_full_pe_raw.csv 2>/dev/null
for k in seed_kernel_pe extend_kernel extend_finish_kernel align_fast_kernel_pe; do python scripts/ncu_lines.py $O/r02_full_pe.ncu-rep "$k" 25 > $O/r02_lines_$k.txt 2>&1; done
rm -f $O/r02_full_pe.ncu-rep
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"xdrop_tile|tail_plan|tail_decide|^align_kernel" -c 40 \
    -o $O/r02_full_cfg5 python tests/tools/run_config.py config5 100000 > $O/r02_full_cfg5.log 2>&1; lap "full cfg5 rc=$?"
ncu -i $O/r02_full_cfg5.ncu-rep --page raw --csv > $O/r02_full_cfg5_raw.csv 2>/dev/null
python scripts/ncu_lines.py $O/r02_full_cfg5.ncu-rep "xdrop_tile_kernel" 25 > $O/r02_lines_xdrop_tile_kernel.txt 2>&1
python scripts/ncu_lines.py $O/r02_full_cfg5.ncu-rep "^align_kernel" 25 > $O/r02_lines_align_kernel.txt 2>&1
rm -f $O/r02_full_cfg5.ncu-rep
timeout ${MEMCHECK_S:-480} compute-sanitizer --tool memcheck --error-exitcode 1 --print-limit 20 python -m pytest tests/test_map_paired_parity.py tests/test_xdrop_golden.py ${MEMCHECK_MORE:-} -m gpu -q \
    > $O/r02_sanitizer_memcheck.log 2>&1; lap "memcheck rc=$?"; tail -3 $O/r02_sanitizer_memcheck.log
timeout ${RACECHECK_S:-360} compute-sanitizer --tool racecheck --error-exitcode 1 --print-limit 20 python -m pytest tests/test_map_paired_parity.py tests/test_xdrop_golden.py -m gpu -q -k "rescue or vectors or repeats" \
    > $O/r02_sanitizer_racecheck.log 2>&1; lap "racecheck rc=$?"; tail -3 $O/r02_sanitizer_racecheck.log
