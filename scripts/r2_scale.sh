#!/bin/bash
# round-2 scaling pass on one 8-GPU box: bench lines at N = 1, 2, 4, 8 back to back (weak, with the strong split, the
# parity objects and the e2e legs inside); FRB_SCALE_TESTS=1 adds the 2 / 4 / 8-rank parity tests
mkdir -p gpurun_out
if [ -n "$FRB_SCALE_TESTS" ]; then
  FRB_REQUIRE_GPUS=8 timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "8] or 4-ssprk3 or rhs[4 or hooks[auto-limiter-4 or hooks[auto-filter-4 or cavity_slabs_match_oracle[4 or eight_rank or host_between_steps[4 or rows[16" > gpurun_out/r2_multi_tests_8.log 2>&1
  tail -4 gpurun_out/r2_multi_tests_8.log
fi
for N in ${FRB_SCALE_NS:-1 2 4 8}; do
  if [ $N = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_scale_$N.json 2> gpurun_out/r2_scale_$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_scale_$N.json 2> gpurun_out/r2_scale_$N.err
  fi
  tail -2 gpurun_out/r2_scale_$N.err | cut -c1-200
done
python - <<PY
import json
for N in (1, 2, 4, 8):
    f = "gpurun_out/r2_scale_%d.json" % N
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        s = d.get("strong") or {}
        print(N, "weak value %.4g ms/step %.4f avg_launch %.4f parity %s e2e %.3g launches %d" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["parity"]["ok"], d["e2e"]["value"], d["gpu_launches"]))
        if s: print(N, "strong value %.4g ms/step %.4f avg_launch %.4f parity %s e2e %.3g" % (s.get("value", 0), s.get("ms_per_step", 0), s.get("avg_launch_ms", 0), (s.get("parity") or {}).get("ok"), (s.get("e2e") or {}).get("value", 0)))
    except Exception as e:
        print(N, "ERR", e)
PY
