#!/bin/bash
# round-2 closing pass after the curvilinear kernels changed: ncu capture of the one-launch kernel, every GPU test,
# smoke, the f2 bench line and a short headline line
mkdir -p gpurun_out
PROBE_ROWS=32 timeout 300 ncu --set full --clock-control none --import-source on -k regex:curv_fused --launch-skip 3 -c 1 \
  -o gpurun_out/r2_curv_fused_final python scripts/probe_curv_march.py 1024 1024 3 2 > gpurun_out/r2_curv_fused_ncu.log 2>&1
tail -2 gpurun_out/r2_curv_fused_ncu.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_tests_final2.log 2>&1
tail -4 gpurun_out/r2_tests_final2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --config f2 --steps 10 --warmup 3 > gpurun_out/r2_bench_cfgf2.json 2> gpurun_out/r2_bench_cfgf2.err
tail -c 1500 gpurun_out/r2_bench_cfgf2.json; tail -3 gpurun_out/r2_bench_cfgf2.err
timeout 200 python scripts/probe_curv.py 2048 1024 2 20 2>&1 | head -2 | tee gpurun_out/r2_curv_p2.jsonl
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2_bench_1d.json 2> gpurun_out/r2_bench_1d.err
tail -c 300 gpurun_out/r2_bench_1d.json; tail -3 gpurun_out/r2_bench_1d.err
