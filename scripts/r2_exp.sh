#!/bin/bash
mkdir -p gpurun_out
V=fluxreconstruction.jl_b200/lib/variants
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -q -k "ns_ or cfg5 or cavity" > gpurun_out/r2_tests5.log 2>&1
tail -4 gpurun_out/r2_tests5.log
echo "== ns"; python scripts/ns_probe.py
for v in default b4n2 b3n2 b3n4; do
  echo "== rc $v"; if [ $v = default ]; then python scripts/probe_cfg3.py 2048 rc 2>&1 | grep stage_kind; else FRB200_LIB=$V/libfrb200_$v.so python scripts/probe_cfg3.py 2048 rc 2>&1 | grep stage_kind; fi
done
echo "== bgk default"; python scripts/bgk_probe.py
echo "== bgk minb3"; FRB200_LIB=$V/libfrb200_minb3.so python scripts/bgk_probe.py
echo "== bgk two-pass"; FRB_BGK_TWO_PASS=1 python scripts/bgk_probe.py
