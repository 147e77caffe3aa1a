#!/bin/bash
# Round-2 GPU pass G (1 GPU): HRNet-W48 plan (teacher-forced pin + throughput), smoke(), W32 bench (fp16 line too).
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_teacher_forced.py -q -s -k "w48 or 16bit" > $OUT/g_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/g_pytest.log
timeout 600 python __graft_entry__.py smoke > $OUT/g_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/g_smoke.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 400 $B --backbone hrnet_w48 > $OUT/g_bench_w48.json 2> $OUT/g_bench.err
timeout 400 $B --dtype fp16 > $OUT/g_bench_fp16.json 2>> $OUT/g_bench.err
timeout 400 $B > $OUT/g_bench_bf16.json 2>> $OUT/g_bench.err
grep -E "teacher-forced|passed|failed|exit" $OUT/g_pytest.log | cut -c1-400; tail -6 $OUT/g_smoke.log
for f in w48 fp16 bf16; do python -c "
import json,sys
try:
    txt=[l for l in open('$OUT/g_bench_$f.json') if l.startswith('{')][-1]
    j=json.loads(txt); print('$f', round(j['value'],1), round(j['ms_per_step'],2), round(j['roofline']['conv_ms_per_step'],2), round(j['roofline']['frac'],3), j['clocks']['sm_mhz'], round(j['roofline']['whole_net_tflops'],1))
except Exception as e: print('$f', 'ERR', e); print(open('$OUT/g_bench.err').read()[-800:])
"; done
