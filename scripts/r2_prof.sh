#!/bin/bash
# round-2 profiling pass (1 GPU): A/B of the row-chunk kernel with / without the slab-parallel code, ncu captures
mkdir -p gpurun_out
V=fluxreconstruction.jl_b200/lib/variants
echo "== default"; python scripts/probe_cfg3.py 2048 rc 2>&1 | grep stage_kind
echo "== nohalo"; FRB200_LIB=$V/libfrb200_nohalo.so python scripts/probe_cfg3.py 2048 rc 2>&1 | grep stage_kind
echo "== default again"; python scripts/probe_cfg3.py 2048 rc 2>&1 | grep stage_kind
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:euler2d_rc_kernel -s 12 -c 3 -f -o gpurun_out/r2_rc python scripts/probe_cfg3.py 2048 rc > gpurun_out/r2_rc_ncu.log 2>&1
$NCU -k regex:bgk1d_fused -s 6 -c 2 -f -o gpurun_out/r2_bgk python scripts/bgk_probe.py > gpurun_out/r2_bgk_ncu.log 2>&1
$NCU -k regex:ns_ -s 18 -c 6 -f -o gpurun_out/r2_ns python scripts/ns_probe.py > gpurun_out/r2_ns_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
python scripts/bgk_probe.py
FRB_BGK_TWO_PASS=1 python scripts/bgk_probe.py
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/r2_bench_cfg5.json 2> gpurun_out/r2_bench_cfg5.err
tail -c 1500 gpurun_out/r2_bench_cfg5.json; tail -3 gpurun_out/r2_bench_cfg5.err
