#!/bin/bash
# Round-2 GPU pass L (1 GPU): one ncu --set full capture of the fused stem kernel.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_tc -s 3 -c 1 -f -o $OUT/r2_stem_full \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/l_ncu.log 2>&1
echo "ncu exit $?"
tail -3 $OUT/l_ncu.log | cut -c1-300
