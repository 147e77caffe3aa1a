#!/bin/bash
# Round-2 GPU pass C (1 GPU): single-box (P1, base offset 0) correctness + A/B, merged head stems, targeted tests,
# same-box A/B of the whole step (bench variants).
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_conv.py -q > $OUT/c_conv_p1.log 2>&1; P1=$?
echo "p1_exit=$P1" | tee $OUT/c_decision.txt
if [ $P1 -ne 0 ]; then export ACR_B200_P1=0; echo "P1 FAILED -> three boxes" | tee -a $OUT/c_decision.txt; fi
S1="64,64,3,1,64,0"; S2="64,64,3,1,128,0,64,4"; S3="64,64,3,1,64,1"; S4="64,64,3,1,128,1,64,4"
LAYERS="$S1 $S2 $S3 $S4 128,128,3,1,32,1 128,128,3,1,32,0 256,256,3,1,16,1 64,64,3,1,128,0 34,256,3,1,128,0 64,33,3,1,256,0 256,32,3,1,128,0"
for CFG in "1 1" "0 1" "1 2" "1 0"; do
  set -- $CFG
  [ $1 -eq 1 ] && [ $P1 -ne 0 ] && continue
  echo "== P1=$1 EPI=$2" >> $OUT/c_conv_ab.log
  ACR_B200_P1=$1 ACR_B200_EPI=$2 timeout 300 python tools/conv_bench.py $LAYERS >> $OUT/c_conv_ab.log 2>&1
done
timeout 1500 python -m pytest tests/test_gpu_teacher_forced.py tests/test_gpu_network.py -q -s \
    -k "teacher_forced_16bit or heads_only or same_rounding or dropin or cuda_graph or full_batch or channel_slice or head_forward or engine_cache or outputs_survive" > $OUT/c_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/c_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
timeout 400 $B > $OUT/c_bench_default.json 2> $OUT/c_bench.err
ACR_B200_MERGE_STEMS=0 timeout 400 $B > $OUT/c_bench_nomerge.json 2>> $OUT/c_bench.err
ACR_B200_EPI=0 timeout 400 $B > $OUT/c_bench_epi0.json 2>> $OUT/c_bench.err
ACR_B200_P1=0 timeout 400 $B > $OUT/c_bench_p1off.json 2>> $OUT/c_bench.err
ACR_B200_LIB=$PWD/arbitrary-hands-3d-reconstruction_b200/lib/libacr_b200_single.so ACR_B200_P1=0 ACR_B200_EPI=0 ACR_B200_MERGE_STEMS=0 timeout 400 $B > $OUT/c_bench_r1like.json 2>> $OUT/c_bench.err
cat $OUT/c_decision.txt; cat $OUT/c_conv_ab.log; tail -4 $OUT/c_pytest.log
for f in default nomerge epi0 p1off r1like; do python -c "
import json,sys
try:
    j=json.load(open('$OUT/c_bench_$f.json')); print('$f', round(j['value'],1), round(j['ms_per_step'],2), round(j['roofline']['conv_ms_per_step'],2), j['clocks']['sm_mhz'])
except Exception as e: print('$f', 'ERR', e)
"; done
