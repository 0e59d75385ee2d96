#!/bin/bash
# compute-sanitizer over the kernels added for the curvilinear path and the kinetic-advection model (run under gpurun):
# memcheck on every small case, racecheck on the element kernel's shared-memory tiles.
# Output: gpurun_out/sanitize_curv_{memcheck,racecheck}.log
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 0 --print-limit 20 \
  python -m pytest tests/test_gpu_curv.py tests/test_gpu_parity.py -m gpu -q -k "(curv or kinetic) and not large" \
  > gpurun_out/sanitize_curv_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/sanitize_curv_memcheck.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 0 --print-limit 20 \
  python -m pytest tests/test_gpu_curv.py -m gpu -q -k "parallelogram_rhs or cylinder_rhs or metric_from_vertices or parallelogram_steps" \
  > gpurun_out/sanitize_curv_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/sanitize_curv_racecheck.log
