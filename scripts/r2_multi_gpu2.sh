set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
PCU_DENSE_TICKS=1 python bench.py --gpus 1 --steps 20 --warmup 12 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
grep "dense phase" gpurun_out/r2f_bench_n1.err | tail -2
python -m pytest tests/test_gpu_gram_paths.py -m gpu -q -x 2>&1 | tail -3
$TR --nproc-per-node 2 --master-port 29515 bench.py --gpus 2 --steps 20 --warmup 12 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
PCU_NO_CHAIN=1 $TR --nproc-per-node 2 --master-port 29516 bench.py --gpus 2 --steps 20 --warmup 12 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2f_bench_n2_nochain.json 2> gpurun_out/r2f_bench_n2_nochain.err
PCU_NO_CHAIN=1 PCU_NO_UPDSTATS=1 $TR --nproc-per-node 2 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 12 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2f_bench_n2_plain.json 2> gpurun_out/r2f_bench_n2_plain.err
