#!/bin/bash
# round-2 multi-GPU validation: parity tests at N ranks, bench lines at N (weak + the strong split inside), legacy A/B
N=${1:-2}
mkdir -p gpurun_out
export FRB_REQUIRE_GPUS=$N
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2_multi_tests_$N.log 2>&1
tail -3 gpurun_out/r2_multi_tests_$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_$N.json 2> gpurun_out/r2_bench_$N.err
tail -c 2500 gpurun_out/r2_bench_$N.json; tail -5 gpurun_out/r2_bench_$N.err
FRB_HALO_LEGACY=1 timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-strong > gpurun_out/r2_bench_${N}_legacy.json 2> gpurun_out/r2_bench_${N}_legacy.err
python - <<PY
import json
for f in ("gpurun_out/r2_bench_$N.json", "gpurun_out/r2_bench_${N}_legacy.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms/step %.4f avg_launch %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"]), d.get("parity"), (d.get("strong") or {}).get("value"), (d.get("strong") or {}).get("ms_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY
