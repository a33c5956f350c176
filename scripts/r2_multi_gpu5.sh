# N-GPU check of the build with the host-side all-reduce of big_fetch and the bench's
# events outside the timed region: parity verdict + strong-scaling bench line (no e2e)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
N=${1:-2}
TAG=${2:-r2p}
timeout 300 $TR --nproc-per-node $N --master-port 29521 scripts/mgpu_check.py > gpurun_out/${TAG}_mgpu$N.log 2>&1; echo "mgpu_check rc=$?"
grep MGPU_VERDICT gpurun_out/${TAG}_mgpu$N.log | cut -c1-700
timeout 300 $TR --nproc-per-node $N --master-port 29523 bench.py --gpus $N --steps 20 --warmup 12 --no-e2e > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench rc=$?"
tail -2 gpurun_out/${TAG}_bench_n$N.err
python scripts/bench_summary.py gpurun_out/${TAG}_bench_n$N.json | head -14
python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench_n$N.json'));print('value',d['value'],'ms',d['ms_per_step'],'parity',d['parity']['first_violation'],d['parity']['worst'], d['roofline'].get('events_over'))"
PCU_SHM_BIG=1 timeout 300 $TR --nproc-per-node $N --master-port 29525 bench.py --gpus $N --steps 20 --warmup 12 --no-e2e --no-parity --no-cpu-baseline > gpurun_out/${TAG}_bench_n${N}_shmbig.json 2> gpurun_out/${TAG}_bench_n${N}_shmbig.err; echo "bench(host-side all-reduce) rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench_n${N}_shmbig.json'));print('host-side all-reduce variant: value',d['value'],'ms',d['ms_per_step'])"
