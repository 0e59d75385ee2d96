#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bgk or ns_ or cfg4 or cfg5 or kinetic or other_fluxes or supersonic" > gpurun_out/r2_tests4.log 2>&1
tail -5 gpurun_out/r2_tests4.log
echo "== bgk fused"; python scripts/bgk_probe.py
echo "== bgk two-pass"; FRB_BGK_TWO_PASS=1 python scripts/bgk_probe.py
echo "== ns"; python scripts/ns_probe.py
echo "== rc"; python scripts/probe_cfg3.py 2048 rc 2>&1 | grep stage_kind
