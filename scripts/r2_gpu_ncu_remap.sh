#!/bin/bash
# round-2: full ncu capture of the remap kernel (C192 L127 x9: 6*192*48 = 55296 column groups) + launch list of the step
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_remap4 -s 2 -c 1 -o gpurun_out/prof_remap4_${TAG:-cur} -f \
  python bench.py --n 192 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_remap4.log 2>&1
tail -2 gpurun_out/ncu_remap4.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 40 -c 14 --csv \
  --log-file gpurun_out/launches_${TAG:-cur}.csv python bench.py --n 192 --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_${TAG:-cur}.csv | python -c "
import csv,sys
for r in csv.reader(sys.stdin):
    if len(r)>10 and r[0]!='ID': print(r[4][:40], r[-3], r[-2], r[-1])
"
