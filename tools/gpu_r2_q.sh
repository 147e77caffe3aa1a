#!/bin/bash
# Round-2 GPU pass Q (1 GPU): ncu --set full of parthead_kernel (where do its 0.34 ms go?).
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:parthead -s 3 -c 1 -f -o $OUT/r2_parthead_full \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/q_ncu.log 2>&1
echo "ncu exit $?"
