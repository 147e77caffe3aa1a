#!/bin/bash
# Round-2 GPU pass D (1 GPU): x-paired stride-2 convs (MODE_S2X) correctness + A/B, staged epilogue for all N=64 layers,
# teacher-forced sweep of the new plan, same-box bench A/B.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_conv.py -q > $OUT/d_conv.log 2>&1; CV=$?
echo "conv_exit=$CV" | tee $OUT/d_decision.txt
grep -E "passed|failed|FAILED" $OUT/d_conv.log | tail -8 >> $OUT/d_decision.txt
if grep -q "FAILED.*stride2_x_paired" $OUT/d_conv.log; then export ACR_B200_S2X=0; echo "S2X FAILED -> nine parity boxes" | tee -a $OUT/d_decision.txt; fi
echo "== nine parity boxes (CK=32 rows)" >> $OUT/d_conv_ab.log
timeout 300 python tools/conv_bench.py 32,32,3,2,128,0 32,64,3,2,128,0 32,128,3,2,64,0 32,256,3,2,32,0 >> $OUT/d_conv_ab.log 2>&1
echo "== x-paired input (MODE_S2X)" >> $OUT/d_conv_ab.log
timeout 300 python tools/conv_bench.py 32,32,3,2,128,0,128,8 32,64,3,2,128,0,128,8 32,128,3,2,64,0,64,8 32,256,3,2,32,0,32,8 >> $OUT/d_conv_ab.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_teacher_forced.py tests/test_gpu_network.py -q -s \
    -k "teacher_forced_16bit or same_rounding or dropin or full_batch" > $OUT/d_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/d_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 400 $B > $OUT/d_bench_default.json 2> $OUT/d_bench.err
ACR_B200_S2X=0 timeout 400 $B > $OUT/d_bench_s2x0.json 2>> $OUT/d_bench.err
ACR_B200_EPI=0 timeout 400 $B > $OUT/d_bench_epi0.json 2>> $OUT/d_bench.err
cat $OUT/d_decision.txt; cat $OUT/d_conv_ab.log; tail -4 $OUT/d_pytest.log
for f in default s2x0 epi0; do python -c "
import json,sys
try:
    j=json.load(open('$OUT/d_bench_$f.json')); print('$f', round(j['value'],1), round(j['ms_per_step'],2), round(j['roofline']['conv_ms_per_step'],2), j['clocks']['sm_mhz'])
except Exception as e: print('$f', 'ERR', e)
"; done
