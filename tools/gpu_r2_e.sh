#!/bin/bash
# Round-2 GPU pass E (1 GPU): N-split (N=256 as two double-buffered halves) correctness + A/B, whole-step A/B.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_conv.py -q > $OUT/e_conv.log 2>&1; CV=$?
echo "conv_exit=$CV" | tee $OUT/e_decision.txt
grep -E "passed|failed|FAILED" $OUT/e_conv.log | tail -8 >> $OUT/e_decision.txt
if [ $CV -ne 0 ]; then export ACR_B200_NSPLIT=0; echo "NSPLIT FAILED -> whole N" | tee -a $OUT/e_decision.txt; fi
LAYERS="34,256,3,1,128,0 34,256,3,2,128,0 64,256,1,1,128,1 64,256,1,1,128,0 256,256,3,1,16,1 256,256,3,1,16,0 128,256,3,2,32,0 64,256,3,2,32,0"
for NS in 1 0; do
  echo "== NSPLIT=$NS" >> $OUT/e_conv_ab.log
  ACR_B200_NSPLIT=$NS timeout 300 python tools/conv_bench.py $LAYERS >> $OUT/e_conv_ab.log 2>&1
done
timeout 1200 python -m pytest tests/test_gpu_teacher_forced.py tests/test_gpu_network.py -q -s \
    -k "teacher_forced_16bit or same_rounding or dropin or full_batch" > $OUT/e_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/e_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 400 $B > $OUT/e_bench_default.json 2> $OUT/e_bench.err
ACR_B200_NSPLIT=0 timeout 400 $B > $OUT/e_bench_nsplit0.json 2>> $OUT/e_bench.err
cat $OUT/e_decision.txt; cat $OUT/e_conv_ab.log; tail -4 $OUT/e_pytest.log
for f in default nsplit0; do python -c "
import json,sys
try:
    j=json.load(open('$OUT/e_bench_$f.json')); print('$f', round(j['value'],1), round(j['ms_per_step'],2), round(j['roofline']['conv_ms_per_step'],2), j['clocks']['sm_mhz'], j['profile_ms_by_kind'])
except Exception as e: print('$f', 'ERR', e)
"; done
