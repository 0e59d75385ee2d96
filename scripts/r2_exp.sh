#!/bin/bash
mkdir -p gpurun_out
V=fluxreconstruction.jl_b200/lib/variants
for v in default uponly default uponly; do
  echo "== rc $v"
  if [ $v = default ]; then unset FRB200_LIB; else export FRB200_LIB=$V/libfrb200_$v.so; fi
  python scripts/probe_cfg3.py 2048 rc 2>&1 | grep stage_kind
  python bench.py --steps 20 --warmup 5 --no-cpu --no-parity 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench value %.4g ms/step %.4f avg_launch %.4f'%(d['value'],d['ms_per_step'],d['roofline']['avg_launch_ms']), d['clocks'])"
done
unset FRB200_LIB
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "euler2d" > gpurun_out/r2_tests7.log 2>&1; tail -3 gpurun_out/r2_tests7.log
echo "== bgk one-pass"; python scripts/bgk_probe.py 8192
echo "== bgk two-pass"; FRB_BGK_TWO_PASS=1 python scripts/bgk_probe.py 8192
