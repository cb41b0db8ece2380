#!/bin/bash
# k_remap3 with the {1/dp2, pe2} ring and lazy fillz sums: parity + quick benches (fp64 C768; fp32 C384 with 4 / 6 CTAs per SM)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "remap or step or map" ) > gpurun_out/pytest_remap.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_remap.log
grep -E "FAILED|passed|failed|rc=" gpurun_out/pytest_remap.log | cut -c1-300 | tail -8
run() { # tag, env..., args
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity $ARGS > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    print("$tag", d["ms_per_step"], {k:v["avg_ms"] if isinstance(v,dict) else v for k,v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("ERR $tag", e); print(open("gpurun_out/bench_$tag.err").read()[-1500:])
P
}
ARGS="" run f64_c768 X=1
ARGS="--n 384 --dtype float32" run f32_c384_mb6 X=1
ARGS="--n 384 --dtype float32" run f32_c384_mb4 FV3T_REMAP_MINB=4
ARGS="--n 384 --dtype float32" run f32_c384_mb5 FV3T_REMAP_MINB=5
