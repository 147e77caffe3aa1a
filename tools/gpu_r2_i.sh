#!/bin/bash
# Round-2 GPU pass I (1 GPU): fuse sums folded into the producing stride-2 conv: teacher-forced pin + same-box A/B.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_conv.py tests/test_gpu_teacher_forced.py tests/test_gpu_network.py -q -s \
    -k "conv_tc or teacher_forced or same_rounding or dropin or full_batch or golden or head_forward" > $OUT/i_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/i_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 400 $B > $OUT/i_bench_default.json 2> $OUT/i_bench.err
ACR_B200_FOLD_FUSE=0 timeout 400 $B > $OUT/i_bench_nofold.json 2>> $OUT/i_bench.err
timeout 400 $B > $OUT/i_bench_default2.json 2>> $OUT/i_bench.err
grep -E "teacher-forced|passed|failed|exit|Error" $OUT/i_pytest.log | cut -c1-300
for f in default nofold default2; do python -c "
import json
try:
    j=json.loads([l for l in open('$OUT/i_bench_$f.json') if l.startswith('{')][-1]); print('$f', round(j['value'],1), round(j['ms_per_step'],2), round(j['roofline']['conv_ms_per_step'],2), j['clocks']['sm_mhz'], j['profile_ms_by_kind'])
except Exception as e: print('$f', 'ERR', e); print(open('$OUT/i_bench.err').read()[-600:])
"; done
