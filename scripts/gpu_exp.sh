mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for G in 3 2 1; do
FV3T_REMAP_G=$G timeout 300 python bench.py --n 384 --npz 127 --steps 3 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('G=$G', d['ms_per_step'], d['roofline']['kernels'])"
done
