#!/bin/bash
# compute-sanitizer passes over small GPU cases of every kernel family (run under gpurun):
#   memcheck  -- out-of-bounds / misaligned global and shared accesses
#   racecheck -- shared-memory hazards of the marching kernels (mbarrier / bulk-copy rings)
# Output: gpurun_out/sanitize_{memcheck,racecheck}.log
mkdir -p gpurun_out
SEL='euler2d_rhs or row_chunk_steps or zero_dt or bgk_rhs or ns_cavity_rhs or limiter or ghost_fill or modal_filter or tableau or other_fluxes or advection_rhs or euler1d_rhs'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 0 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_tri.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/sanitize_memcheck.log; tail -4 gpurun_out/sanitize_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 0 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "euler2d_rhs or zero_dt or pipelined" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -c "hazard" gpurun_out/sanitize_racecheck.log; tail -4 gpurun_out/sanitize_racecheck.log
