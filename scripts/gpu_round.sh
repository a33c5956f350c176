#!/bin/bash
# One GPU-box pass: parity tests, bench (both arms), ncu launch list and one ncu --set full capture.
# usage (through gpurun): bash scripts/gpu_round.sh <tag>
tag=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
python bench.py --steps 20 --warmup 12 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python scripts/bench_summary.py gpurun_out/bench_$tag.json
python bench.py --impl reference --steps 6 --warmup 12 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_$tag.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 2 --warmup 12 --no-e2e --no-cpu-baseline > gpurun_out/launches_$tag.log 2>&1; echo "ncu list rc=$?"
# iteration 14 of the bench workload (quasi-Newton memory full), bracketed by cudaProfilerStart/Stop
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -o gpurun_out/prof_$tag -f python scripts/profile_run.py --config C3 --n 67108864 --iters 13 --capture 1 > gpurun_out/prof_$tag.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -8
