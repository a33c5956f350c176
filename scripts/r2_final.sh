#!/bin/bash
# Final pass of round 2 on one GPU: full parity suite, smoke(), the default bench line, C2 / C4
# lines, ncu launch list of the bench command and one --set full capture of a C3 iteration
# (condensed on the box).   usage (through gpurun): bash scripts/r2_final.sh <tag>
tag=${1:-r2w}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_$tag.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 12 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python scripts/bench_summary.py gpurun_out/bench_$tag.json | head -3
python -c "
import json;d=json.load(open('gpurun_out/bench_$tag.json'));print('e2e',d['e2e']['value'],'cpu',d['cpu_baseline']['value'],'frac',d['roofline']['frac'],'parity',d['parity']['first_violation'],d['parity']['worst'],d['clocks'])"
for c in C2 C4; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 12 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/bench_${c}_$tag.json 2> gpurun_out/bench_${c}_$tag.err; echo "bench $c rc=$?"
  python scripts/bench_summary.py gpurun_out/bench_${c}_$tag.json | head -4
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 2 --warmup 12 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/launches_$tag.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -o /tmp/prof_$tag -f python scripts/profile_run.py --config C3 --n 67108864 --iters 13 --capture 1 > gpurun_out/prof_$tag.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_summary.py /tmp/prof_$tag.ncu-rep gpurun_out/ncu_summary_$tag.md gpurun_out/traffic_$tag.json > /dev/null; echo "summary rc=$?"
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv --print-units base > gpurun_out/ncu_raw_$tag.csv 2>/dev/null
cat gpurun_out/ncu_summary_$tag.md
