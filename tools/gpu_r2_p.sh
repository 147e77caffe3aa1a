#!/bin/bash
# Round-2 GPU pass P (1 GPU): evidence of the final build after the tcgen05 pooling kernel -- full gpu test suite, launch
# list of one step, ncu --set full of the stem and pooling kernels, default bench line.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q > $OUT/p_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/p_pytest.log
LIST_ONLY=1 bash tools/ncu_conv.sh r2_final > $OUT/p_ncu.log 2>&1
python tools/layer_table.py kernels $OUT/launches_r2_final.csv > $OUT/p_kernels.md
for K in stem_tc pool_tc; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $OUT/r2_${K}_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/p_ncu_$K.log 2>&1
done
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/p_bench_default.json 2> $OUT/p_bench.err
tail -3 $OUT/p_pytest.log; cat $OUT/p_kernels.md | head -24
python -c "
import json
j=json.loads([l for l in open('$OUT/p_bench_default.json') if l.startswith('{')][-1]); print('default', round(j['value'],1), round(j['ms_per_step'],2), round(j['e2e']['value'],1), round(j['roofline']['frac'],4), j['roofline']['traffic'], j['clocks'], j['cpu_baseline']['value'], j['profile_ms_by_kind'])"
