#!/bin/bash
# round-2 multi-GPU validation: parity tests at N ranks, then bench lines at N = 1 and N
N=${1:-2}
mkdir -p gpurun_out
export FRB_REQUIRE_GPUS=$N
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2_multi_tests_$N.log 2>&1
tail -3 gpurun_out/r2_multi_tests_$N.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_bench_1.json 2> gpurun_out/r2_bench_1.err
tail -c 600 gpurun_out/r2_bench_1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_$N.json 2> gpurun_out/r2_bench_$N.err
tail -c 1500 gpurun_out/r2_bench_$N.json; tail -5 gpurun_out/r2_bench_$N.err
