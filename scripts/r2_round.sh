#!/bin/bash
# Round-2 GPU-box pass: parity tests, bench (both arms), C2 / C4 bench lines, ncu launch list
# and one ncu --set full capture of a steady-state iteration.
# usage (through gpurun): bash scripts/r2_round.sh <tag> [skip-tests] [skip-ref]
tag=${1:-r2}
mkdir -p gpurun_out
if [ "$2" != "skip-tests" ]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
fi
python bench.py --steps 20 --warmup 12 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python scripts/bench_summary.py gpurun_out/bench_$tag.json
for c in C2 C4; do
  python bench.py --config $c --steps 10 --warmup 12 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/bench_${c}_$tag.json 2> gpurun_out/bench_${c}_$tag.err; echo "bench $c rc=$?"
  python scripts/bench_summary.py gpurun_out/bench_${c}_$tag.json
done
if [ "$3" != "skip-ref" ]; then
  python bench.py --impl reference --steps 4 --warmup 12 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_$tag.json
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 2 --warmup 12 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/launches_$tag.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -o gpurun_out/prof_$tag -f python scripts/profile_run.py --config C3 --n 67108864 --iters 13 --capture 1 > gpurun_out/prof_$tag.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -o gpurun_out/prof_C4_$tag -f python scripts/profile_run.py --config C4 --n 33554432 --iters 22 --capture 1 > gpurun_out/prof_C4_$tag.log 2>&1; echo "ncu full C4 rc=$?"
ls -la gpurun_out | tail -12
