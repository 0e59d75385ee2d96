#!/bin/bash
# round-2 single-GPU pass: parity tests, smoke, the headline bench line (both arms), the per-config lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_tests_final.log 2>&1
tail -4 gpurun_out/r2_tests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_1.json 2> gpurun_out/r2_bench_1.err
tail -c 400 gpurun_out/r2_bench_1.json; tail -3 gpurun_out/r2_bench_1.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
tail -c 700 gpurun_out/r2_bench_ref.json; tail -3 gpurun_out/r2_bench_ref.err
for c in 1 2 4 5 f2; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2_bench_cfg$c.json 2> gpurun_out/r2_bench_cfg$c.err
  tail -c 200 gpurun_out/r2_bench_cfg$c.json; tail -3 gpurun_out/r2_bench_cfg$c.err
done
