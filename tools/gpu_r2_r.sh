#!/bin/bash
# Round-2 GPU pass R (1 GPU): part head with coalesced chunk merge: parity (network + teacher-forced tests) and step time.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_teacher_forced.py -q -k "not fp32 and not w48" > $OUT/r_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/r_pytest.log
tail -3 $OUT/r_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
for f in a b; do
timeout 400 $B > $OUT/r_bench_$f.json 2> $OUT/r_bench.err
python -c "
import json
try:
    j=json.loads([l for l in open('$OUT/r_bench_$f.json') if l.startswith('{')][-1]); print('$f', round(j['value'],1), round(j['ms_per_step'],2), j['clocks']['sm_mhz'], j['profile_ms_by_kind'])
except Exception as e: print('$f', 'ERR', e); print(open('$OUT/r_bench.err').read()[-600:])
"; done
