#!/bin/bash
# Round-2 GPU pass B (1 GPU): UMMA unaligned-start probe, ring-staged epilogue correctness + A/B, the tests pass A did not
# reach (it ran with -x), default bench.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 120 tools/umma_shift_probe > $OUT/b_probe.log 2>&1; echo "probe exit $?" >> $OUT/b_probe.log
timeout 600 python -m pytest tests/test_gpu_conv.py -q > $OUT/b_conv_epi1.log 2>&1; E1=$?
ACR_B200_EPI=2 timeout 600 python -m pytest tests/test_gpu_conv.py -q > $OUT/b_conv_epi2.log 2>&1; E2=$?
echo "epi1_exit=$E1 epi2_exit=$E2" | tee $OUT/b_decision.txt
if [ $E1 -ne 0 ]; then export ACR_B200_EPI=0; echo "ring-staged epilogue FAILED -> EPI=0" | tee -a $OUT/b_decision.txt; fi
S1="64,64,3,1,64,0"; S2="64,64,3,1,128,0,64,4"; S3="64,64,3,1,64,1"; S4="64,64,3,1,128,1,64,4"
LAYERS="$S1 $S2 $S3 $S4 64,256,1,1,128,1 64,256,1,1,128,0 256,64,1,1,128,0 128,128,3,1,32,1 128,128,3,1,32,0 256,256,3,1,16,1 256,256,3,1,16,0 64,64,3,1,128,0 64,128,3,2,64,0"
for E in 0 1 2; do
  [ $E -eq 1 ] && [ $E1 -ne 0 ] && continue
  [ $E -eq 2 ] && [ $E2 -ne 0 ] && continue
  echo "== EPI=$E" >> $OUT/b_conv_ab.log
  ACR_B200_EPI=$E timeout 300 python tools/conv_bench.py $LAYERS >> $OUT/b_conv_ab.log 2>&1
done
timeout 1500 python -m pytest tests/test_gpu_network.py tests/test_gpu_parse.py tests/test_gpu_preprocess.py tests/test_gpu_teacher_forced.py -q -s \
    -k "channel_slice or refconv or full_batch or dropin or cuda_graph or parse or preprocess or teacher or heads_only" > $OUT/b_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/b_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/b_bench.json 2> $OUT/b_bench.err
cat $OUT/b_probe.log | head -60; cat $OUT/b_decision.txt; cat $OUT/b_conv_ab.log; tail -5 $OUT/b_pytest.log
