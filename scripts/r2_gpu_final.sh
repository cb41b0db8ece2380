#!/bin/bash
# round-2 final 1-GPU pass: the whole GPU test suite, smoke(), then the evidence pass (scripts/r2_gpu_evidence.sh)
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu_full.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_full.log
grep -E "FAILED|passed|failed|rc=|real" gpurun_out/pytest_gpu_full.log | cut -c1-300 | tail -10
( timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
bash scripts/r2_gpu_evidence.sh
