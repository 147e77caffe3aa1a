#!/bin/bash
# Timing floors of the conv kernel's warp roles on the two N=64 layer classes (run through gpurun).
S1="64,64,3,1,64,0"; S2="64,64,3,1,128,0,64,4"; S3="64,64,3,1,64,1"; S4="64,64,3,1,128,1,64,4"
echo "== product kernels"; python tools/conv_bench.py $S1 $S2 $S3 $S4 "128,128,3,1,32,1" "256,256,3,1,16,1" "64,256,1,1,128,1" "64,64,1,1,128,0"
echo "== TMA-store epilogue"; ACR_B200_TMA_OUT=1 python tools/conv_bench.py $S1 $S2 $S3 $S4 "64,256,1,1,128,1"
for D in 1 2 4 3; do echo "== ACR_B200_CONV_DIAG=$D"; ACR_B200_CONV_DIAG=$D python tools/conv_bench.py $S1 $S2 $S3 $S4; done
