#!/bin/bash
# Round-2 GPU pass N (1 GPU): attention pooling on tcgen05 / TMA (csrc/pool_tc.cu): parity + same-box A/B against mma.sync.
set -u
OUT=gpurun_out
mkdir -p $OUT
ACR_B200_DEBUG_SYNC=1 timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_teacher_forced.py -q -s -k "pool or teacher_forced_16bit or same_rounding or dropin" > $OUT/n_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/n_pytest.log
grep -E "teacher-forced|passed|failed|exit|Error|error|timeout|pool" $OUT/n_pytest.log | cut -c1-250 | head -30
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 400 $B > $OUT/n_bench_tc.json 2> $OUT/n_bench.err
ACR_B200_POOL_TC=0 timeout 400 $B > $OUT/n_bench_mmasync.json 2>> $OUT/n_bench.err
timeout 400 $B > $OUT/n_bench_tc2.json 2>> $OUT/n_bench.err
for f in tc mmasync tc2; do python -c "
import json
try:
    j=json.loads([l for l in open('$OUT/n_bench_$f.json') if l.startswith('{')][-1]); print('$f', round(j['value'],1), round(j['ms_per_step'],2), j['clocks']['sm_mhz'], j['profile_ms_by_kind'])
except Exception as e: print('$f', 'ERR', e); print(open('$OUT/n_bench.err').read()[-600:])
"; done
