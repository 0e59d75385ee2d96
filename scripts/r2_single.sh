#!/bin/bash
# round-2 single-GPU pass: parity tests, the headline bench line, the per-config lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_tests3.log 2>&1
tail -4 gpurun_out/r2_tests3.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_1.json 2> gpurun_out/r2_bench_1.err
tail -c 2500 gpurun_out/r2_bench_1.json; tail -3 gpurun_out/r2_bench_1.err
for c in 1 2 4 5 f2; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2_bench_cfg$c.json 2> gpurun_out/r2_bench_cfg$c.err
  tail -c 1800 gpurun_out/r2_bench_cfg$c.json; tail -3 gpurun_out/r2_bench_cfg$c.err
done
