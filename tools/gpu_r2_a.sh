#!/bin/bash
# Round-2 GPU pass A (1 GPU): dual-issuer + single-box (P1) conv correctness with fall-backs, per-layer A/B timings,
# all gpu tests, default bench, MANO stand-alone bench + ncu capture of the MANO kernel.
set -u
OUT=gpurun_out
mkdir -p $OUT
LIBDIR=arbitrary-hands-3d-reconstruction_b200/lib
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/a_smi.txt 2>&1
# ---- 0. conv correctness: dual issuer alone (three boxes, direct epilogue), then + single box, then + staged epilogue
export ACR_B200_EPI=0
ACR_B200_P1=0 timeout 600 python -m pytest tests/test_gpu_conv.py -q -x > $OUT/a_conv_dual_p0.log 2>&1; DUAL_OK=$?
if [ $DUAL_OK -ne 0 ]; then export ACR_B200_LIB=$PWD/$LIBDIR/libacr_b200_single.so; echo "DUAL ISSUER FAILED -> single" | tee $OUT/a_decision.txt; fi
ACR_B200_P1=1 timeout 600 python -m pytest tests/test_gpu_conv.py -q -x > $OUT/a_conv_p1.log 2>&1; P1_OK=$?
if [ $P1_OK -ne 0 ]; then
  ACR_B200_CONV_DIAG=16 timeout 600 python -m pytest tests/test_gpu_conv.py -q -x > $OUT/a_conv_p1_nobo.log 2>&1
  echo "P1 (base offset = kx) FAILED; zero base offset exit $?" | tee -a $OUT/a_decision.txt
  export ACR_B200_P1=0
fi
ACR_B200_EPI=1 timeout 600 python -m pytest tests/test_gpu_conv.py -q -x > $OUT/a_conv_epi.log 2>&1; EPI_OK=$?
if [ $EPI_OK -eq 0 ]; then export ACR_B200_EPI=1; else echo "STAGED EPILOGUE FAILED" | tee -a $OUT/a_decision.txt; fi
echo "dual_ok=$DUAL_OK p1_ok=$P1_OK epi_ok=$EPI_OK lib=${ACR_B200_LIB:-default} P1=${ACR_B200_P1:-default} EPI=$ACR_B200_EPI" | tee -a $OUT/a_decision.txt
# ---- 1. per-layer A/B
S1="64,64,3,1,64,0"; S2="64,64,3,1,128,0,64,4"; S3="64,64,3,1,64,1"; S4="64,64,3,1,128,1,64,4"
LAYERS="$S1 $S2 $S3 $S4 128,128,3,1,32,1 256,256,3,1,16,1 64,64,3,1,128,0 64,256,1,1,128,1"
for LIB in default single; do
  for P in 1 0; do
    for E in 1 0; do
      [ $P -eq 1 ] && [ $P1_OK -ne 0 ] && continue
      [ $E -eq 1 ] && [ $EPI_OK -ne 0 ] && continue
      [ $LIB = default ] && [ $DUAL_OK -ne 0 ] && continue
      [ $LIB = single ] && [ $P$E != 00 ] && [ $P$E != 11 ] && continue
      echo "== lib=$LIB P1=$P EPI=$E" >> $OUT/a_conv_ab.log
      if [ $LIB = single ]; then L=$PWD/$LIBDIR/libacr_b200_single.so; else L=$PWD/$LIBDIR/libacr_b200.so; fi
      ACR_B200_LIB=$L ACR_B200_P1=$P ACR_B200_EPI=$E timeout 300 python tools/conv_bench.py $LAYERS >> $OUT/a_conv_ab.log 2>&1
    done
  done
done
# ---- 2. everything
timeout 1800 python -m pytest tests -m gpu -x -q -s > $OUT/a_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/a_bench.json 2> $OUT/a_bench.err
timeout 300 python tools/mano_bench.py > $OUT/a_mano_bench.jsonl 2>&1
for N in 512 65536; do
  MANO_BENCH_N=$N timeout 600 ncu --set full --import-source on --clock-control none -k regex:mano_forward -s 5 -c 1 -f \
      -o $OUT/mano_r2_$N python tools/mano_bench.py > $OUT/a_mano_ncu_$N.log 2>&1
done
cat $OUT/a_decision.txt; cat $OUT/a_conv_ab.log
tail -5 $OUT/a_pytest.log
tail -c 600 $OUT/a_bench.json
