#!/bin/bash
# ncu evidence with the reports condensed ON the box (a .ncu-rep of a full iteration is
# 60-200 MB; gpurun_out/ carries 64 MB): launch list + --set full of one steady-state
# iteration of C3, launch list + --set full of the Gram / pass-2 kernels of C4.
# usage (through gpurun): bash scripts/r2_profile.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 2 --warmup 12 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/launches_$tag.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -o /tmp/prof_$tag -f python scripts/profile_run.py --config C3 --n 67108864 --iters 13 --capture 1 > gpurun_out/prof_$tag.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_summary.py /tmp/prof_$tag.ncu-rep gpurun_out/ncu_summary_$tag.md gpurun_out/traffic_$tag.json; echo "summary rc=$?"
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv --print-units base > gpurun_out/ncu_raw_$tag.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_C4_$tag.csv \
  python bench.py --config C4 --steps 1 --warmup 12 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/launches_C4_$tag.log 2>&1; echo "ncu list C4 rc=$?"
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:'gram|Pass2|mdot' -c 8 \
  -o /tmp/prof_C4_$tag -f python scripts/profile_run.py --config C4 --n 33554432 --iters 22 --capture 1 > gpurun_out/prof_C4_$tag.log 2>&1; echo "ncu full C4 rc=$?"
python scripts/ncu_summary.py /tmp/prof_C4_$tag.ncu-rep gpurun_out/ncu_summary_C4_$tag.md gpurun_out/traffic_C4_$tag.json; echo "summary C4 rc=$?"
python scripts/measure_fp64_peak.py > gpurun_out/fp64_peak_$tag.json 2> gpurun_out/fp64_peak_$tag.err; cat gpurun_out/fp64_peak_$tag.json
ls -la gpurun_out | tail -14
