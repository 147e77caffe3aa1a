#!/bin/bash
# Round-2 GPU pass J (1 GPU): stem conv with the im2col operand built in shared memory: parity + same-box A/B.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_teacher_forced.py -q -s -k "stem or teacher_forced_16bit or full_batch or dropin or same_rounding" > $OUT/j_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/j_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 400 $B > $OUT/j_bench_default.json 2> $OUT/j_bench.err
ACR_B200_STEM_FUSED=0 timeout 400 $B > $OUT/j_bench_im2col.json 2>> $OUT/j_bench.err
timeout 400 $B > $OUT/j_bench_default2.json 2>> $OUT/j_bench.err
grep -E "teacher-forced|passed|failed|exit|Error|timeout" $OUT/j_pytest.log | cut -c1-250
for f in default im2col default2; do python -c "
import json
try:
    j=json.loads([l for l in open('$OUT/j_bench_$f.json') if l.startswith('{')][-1]); print('$f', round(j['value'],1), round(j['ms_per_step'],2), j['clocks']['sm_mhz'], j['profile_ms_by_kind'])
except Exception as e: print('$f', 'ERR', e); print(open('$OUT/j_bench.err').read()[-600:])
"; done
