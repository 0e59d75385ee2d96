#!/bin/bash
# round-2 single-GPU pass: parity tests, the headline bench line, the per-config lines
mkdir -p gpurun_out
echo "== bgk one-pass"; python scripts/bgk_probe.py 8192
echo "== bgk two-pass"; FRB_BGK_TWO_PASS=1 python scripts/bgk_probe.py 8192
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_tests6.log 2>&1
tail -4 gpurun_out/r2_tests6.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_1.json 2> gpurun_out/r2_bench_1.err
tail -c 600 gpurun_out/r2_bench_1.json; tail -3 gpurun_out/r2_bench_1.err
for c in 1 2 4 5 f2; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2_bench_cfg$c.json 2> gpurun_out/r2_bench_cfg$c.err
  tail -c 300 gpurun_out/r2_bench_cfg$c.json; tail -3 gpurun_out/r2_bench_cfg$c.err
done
