#!/bin/bash
# Round-2 closing pass (1 GPU): full gpu test suite, launch list of one step, default bench line (with the CPU arm).
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q > $OUT/s_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/s_pytest.log
LIST_ONLY=1 bash tools/ncu_conv.sh r2_final > $OUT/s_ncu.log 2>&1
python tools/layer_table.py kernels $OUT/launches_r2_final.csv > $OUT/s_kernels.md
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/s_bench_default.json 2> $OUT/s_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>> $OUT/s_bench.err | grep "^{" | tail -1 > $OUT/s_bench_reference.json
tail -3 $OUT/s_pytest.log; grep -E "pool|stem|parthead|total" $OUT/s_kernels.md
python -c "
import json
j=json.loads([l for l in open('$OUT/s_bench_default.json') if l.startswith('{')][-1]); print('default', round(j['value'],1), round(j['ms_per_step'],2), round(j['e2e']['value'],1), round(j['roofline']['frac'],4), j['roofline']['traffic'], j['clocks'], j['cpu_baseline']['value'], j['profile_ms_by_kind'])
j=json.loads(open('$OUT/s_bench_reference.json').read()); print('reference', round(j['value'],2), j['cpu_baseline']['cores'], j['cpu_baseline']['kind'])"
