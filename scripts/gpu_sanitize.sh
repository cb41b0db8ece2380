#!/bin/bash
# compute-sanitizer passes over small parity cases: memcheck (global / shared / local out-of-bounds) and racecheck (shared-memory hazards
# between the barrier-separated phases and the cp.async ring)
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
  -k "test_tracer_step_equals_the_two_separate_calls or (test_tracer_2d_parity_c24 and fast and float64 and (8 or 13)) or test_remap_prepare" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitize_memcheck.log
tail -6 gpurun_out/sanitize_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
  -k "test_tracer_2d_subcycling and fast and float64 and 1.8" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitize_racecheck.log
tail -6 gpurun_out/sanitize_racecheck.log
