#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/test_gpu_configs.py -m gpu -q -s --durations=10 ) > gpurun_out/pytest_configs.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_configs.log
grep -E "FAILED|passed|failed|rc=|^E  |config [0-9]|real|s call" gpurun_out/pytest_configs.log | cut -c1-400 | tail -40
