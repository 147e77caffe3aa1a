#!/bin/bash
# Round-2 GPU pass M (1 GPU): bench lines of the final build on one box -- default (with the CPU arm and the stamped conv
# traffic), fp16, HRNet-W48, smoke().
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/m_bench_default.json 2> $OUT/m_bench.err
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 400 $B --dtype fp16 > $OUT/m_bench_fp16.json 2>> $OUT/m_bench.err
timeout 400 $B --backbone hrnet_w48 > $OUT/m_bench_w48.json 2>> $OUT/m_bench.err
timeout 400 $B > $OUT/m_bench_default2.json 2>> $OUT/m_bench.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/m_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/m_smoke.log
tail -3 $OUT/m_smoke.log
for f in default fp16 w48 default2; do python -c "
import json
try:
    j=json.loads([l for l in open('$OUT/m_bench_$f.json') if l.startswith('{')][-1]); print('$f', round(j['value'],1), round(j['ms_per_step'],2), round(j['e2e']['value'],1), round(j['roofline']['frac'],4), j['roofline']['traffic'], j['clocks']['sm_mhz'], j['profile_ms_by_kind'])
except Exception as e: print('$f', 'ERR', e); print(open('$OUT/m_bench.err').read()[-600:])
"; done
