set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29511 scripts/mgpu_check.py > gpurun_out/r2e_mgpu8.log 2>&1
$TR --nproc-per-node 2 --master-port 29512 scripts/mgpu_check.py > gpurun_out/r2e_mgpu2.log 2>&1
$TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 12 --no-e2e --no-cpu-baseline > gpurun_out/r2e_bench_n8.json 2> gpurun_out/r2e_bench_n8.err
PCU_NO_CHAIN=1 $TR --nproc-per-node 8 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 12 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2e_bench_n8_nochain.json 2> gpurun_out/r2e_bench_n8_nochain.err
$TR --nproc-per-node 2 --master-port 29515 bench.py --gpus 2 --steps 20 --warmup 12 --no-e2e --no-cpu-baseline > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err
python bench.py --gpus 1 --steps 20 --warmup 12 --no-e2e --no-cpu-baseline > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err
grep MGPU_VERDICT gpurun_out/r2e_mgpu8.log gpurun_out/r2e_mgpu2.log | cut -c1-600
