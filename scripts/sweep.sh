#!/bin/bash
# usage: scripts/sweep.sh "ENV1=a ENV2=b" "ENV1=c" ...   (runs the device-only bench per setting)
for cfg in "$@"; do
  out=gpurun_out/sweep_$(echo "$cfg" | tr ' =/' '___').json
  env $cfg python bench.py --steps 12 --no-e2e --no-cpu-baseline > $out 2>/dev/null
  echo "== $cfg"
  python scripts/bench_summary.py $out 2>/dev/null | head -9
done
