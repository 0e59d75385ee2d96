#!/bin/bash
# compute-sanitizer over the kernels added / changed in round 2 (run under gpurun, 1 GPU; the 2-rank pass needs 2):
#   row-chunk kernel (rhs_only route, downward-marching segments, LF / Roe instantiations), BGK one-pass kernel
#   (cp.async staging), GKS kernels; memcheck + racecheck.  With 2 GPUs: memcheck of a 2-rank slab run (the in-kernel
#   exchange: ring slots in peer memory, mailbox polls).
mkdir -p gpurun_out
SEL='euler2d_rhs or resident_row_chunk or f_is_pure or other_fluxes or supersonic or bgk_rhs or bgk_steps or kinetic or ns_cavity_rhs or row_chunk_steps'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 0 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/r2_sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r2_sanitize_memcheck.log; tail -4 gpurun_out/r2_sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 0 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "euler2d_rhs or resident_row_chunk or bgk_rhs or other_fluxes" > gpurun_out/r2_sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -c "hazard" gpurun_out/r2_sanitize_racecheck.log; tail -4 gpurun_out/r2_sanitize_racecheck.log
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout 600 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 0 --print-limit 20 \
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
    tests/dist/check_dist.py 64 48 4 ssprk3 auto wave_x > gpurun_out/r2_sanitize_memcheck_2rank.log 2>&1
  echo "2-rank memcheck rc=$?"; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r2_sanitize_memcheck_2rank.log; grep "check_dist\|ERROR SUMMARY" gpurun_out/r2_sanitize_memcheck_2rank.log | tail -4
fi
