#!/bin/bash
# Round-2 GPU pass O (1 GPU): one ncu --set full capture of the tcgen05 attention-pooling kernel.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pool_tc -s 3 -c 1 -f -o $OUT/r2_pool_tc_full \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/o_ncu.log 2>&1
echo "ncu exit $?"
