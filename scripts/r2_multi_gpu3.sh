set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="bench.py --gpus 8 --steps 20 --warmup 12 --no-e2e --no-cpu-baseline"
$TR --nproc-per-node 8 --master-port 29511 scripts/mgpu_check.py > gpurun_out/r2j_mgpu8.log 2>&1
$TR --nproc-per-node 8 --master-port 29513 $B > gpurun_out/r2j_bench_n8.json 2> gpurun_out/r2j_bench_n8.err
PCU_NO_CHAIN=1 $TR --nproc-per-node 8 --master-port 29514 $B --no-parity > gpurun_out/r2j_bench_n8_nochain.json 2> gpurun_out/r2j_bench_n8_nochain.err
PCU_NO_SHM=1 $TR --nproc-per-node 8 --master-port 29515 $B --no-parity > gpurun_out/r2j_bench_n8_noshm.json 2> gpurun_out/r2j_bench_n8_noshm.err
PCU_NO_SHM=1 PCU_NO_CHAIN=1 $TR --nproc-per-node 8 --master-port 29516 $B --no-parity > gpurun_out/r2j_bench_n8_noshm_nochain.json 2> gpurun_out/r2j_bench_n8_noshm_nochain.err
$TR --nproc-per-node 4 --master-port 29517 bench.py --gpus 4 --steps 20 --warmup 12 --no-e2e --no-cpu-baseline > gpurun_out/r2j_bench_n4.json 2> gpurun_out/r2j_bench_n4.err
python bench.py --gpus 1 --steps 20 --warmup 12 --no-e2e --no-cpu-baseline > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err
grep MGPU_VERDICT gpurun_out/r2j_mgpu8.log | cut -c1-400
