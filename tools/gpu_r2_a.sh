#!/bin/bash
# Round-2 GPU pass A (1 GPU): all gpu tests, default bench, MANO stand-alone bench + ncu capture of the MANO kernel.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/a_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/a_bench.json 2> $OUT/a_bench.err
timeout 300 python tools/mano_bench.py > $OUT/a_mano_bench.jsonl 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:mano_forward -s 6 -c 2 -f -o $OUT/mano_r2 \
    python tools/mano_bench.py > $OUT/a_mano_ncu.log 2>&1
tail -5 $OUT/a_pytest.log
tail -c 600 $OUT/a_bench.json
