#!/bin/bash
# ncu evidence for the conv kernel (run on the GPU box through gpurun; outputs under gpurun_out/):
#   1. launch list of one whole step (gpu__time_duration.sum per launch)
#   2. --set full + source of one x-paired 32->32 launch (MODE 7) and one 64->64 patch/resident launch (MODE 3)
set -u
TAG=${1:-r1_v13}
OUT=gpurun_out
mkdir -p $OUT
L=$(python tools/profile_step.py --batch 256 --steps 1 | grep "launches per step" | awk '{print $4}')
echo "launches per step: $L"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/launches_${TAG}.csv python tools/profile_step.py --batch 256 --range > $OUT/prof_list.log 2>&1
[ "${LIST_ONLY:-0}" = "1" ] && exit 0
NC=$(python -c "
import sys; sys.path.insert(0, 'arbitrary-hands-3d-reconstruction_b200')
from acr_b200 import lib as L; from acr_b200.engine import Engine
print(sum(1 for r in Engine(None, 1, 'cpu', dry_run=True).recs if r['kind'] == L.OP_CONV))")
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:conv_tc -s "$NC" -c "$NC" --csv --log-file $OUT/conv_traffic_${TAG}.csv \
    python tools/profile_step.py --batch 256 > $OUT/prof_traffic.log 2>&1
python tools/conv_traffic.py $OUT/conv_traffic_${TAG}.csv $OUT/conv_traffic_${TAG}.json
[ "${FULL:-1}" = "1" ] || exit 0
for MODE in 23 19 34; do
  timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled \
      -k "regex:conv_tc_kernel<\(int\)64, __nv_bfloat16, \(int\)${MODE}>" -s 12 -c 2 -f \
      -o $OUT/conv_tc_${TAG}_mode${MODE} python tools/profile_step.py --batch 256 > $OUT/prof_full_${MODE}.log 2>&1
done
ls -la $OUT | tail -5
