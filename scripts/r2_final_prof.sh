#!/bin/bash
# round-2 final profiling pass (1 GPU): ncu captures of the final kernels, launch list of the bench command
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:euler2d_rc_kernel -s 12 -c 3 -f -o gpurun_out/r2_rc_final python scripts/probe_cfg3.py 2048 rc > gpurun_out/r2_rc_final_ncu.log 2>&1
$NCU -k regex:ns_ -s 18 -c 6 -f -o gpurun_out/r2_ns_final python scripts/ns_probe.py > gpurun_out/r2_ns_final_ncu.log 2>&1
$NCU -k regex:bgk -s 9 -c 4 -f -o gpurun_out/r2_bgk_final python scripts/bgk_probe.py > gpurun_out/r2_bgk_final_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_launches_bench.log 2>&1
ls -la gpurun_out/*final*.ncu-rep gpurun_out/r2_launches.csv
python tests/harness/probe_tri.py 192 192 2
python tests/harness/probe_tri.py 128 128 3
python scripts/probe_curv.py 1024 1024 3 10 2>&1 | head -2
