# 2-GPU check of the final round-2 build: golden comparisons partitioned over the ranks
# (scripts/mgpu_check.py: MGPU_VERDICT) + the strong-scaling bench line with its parity block
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
N=${1:-2}
$TR --nproc-per-node $N --master-port 29521 scripts/mgpu_check.py > gpurun_out/r2h_mgpu$N.log 2>&1; echo "mgpu_check rc=$?"
grep MGPU_VERDICT gpurun_out/r2h_mgpu$N.log | cut -c1-600
$TR --nproc-per-node $N --master-port 29523 bench.py --gpus $N --steps 20 --warmup 12 > gpurun_out/r2h_bench_n$N.json 2> gpurun_out/r2h_bench_n$N.err; echo "bench rc=$?"
python scripts/bench_summary.py gpurun_out/r2h_bench_n$N.json | head -4
python -c "
import json;d=json.load(open('gpurun_out/r2h_bench_n$N.json'));print('value',d['value'],'e2e',d['e2e'] and d['e2e']['value'],'parity',d['parity']['first_violation'],d['parity']['worst'])"
