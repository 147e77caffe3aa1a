#!/bin/bash
# Round-end evidence run on one B200 (through gpurun): ncu launch list + conv DRAM traffic, the whole GPU test
# suite, smoke(), the default bench, the reference arm, and the batch-1 latency.  Logs land in gpurun_out/.
TAG=${1:-r1_final}
OUT=gpurun_out; mkdir -p $OUT
FULL=0 bash tools/ncu_conv.sh $TAG > $OUT/ncu_${TAG}.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $OUT/tests_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $OUT/smoke_${TAG}.log
timeout 600 python bench.py 2>/dev/null | tail -1 > $OUT/bench_${TAG}.json; cut -c1-400 $OUT/bench_${TAG}.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > $OUT/bench_ref_${TAG}.json; cut -c1-600 $OUT/bench_ref_${TAG}.json
timeout 300 python tools/latency_bench.py 2>&1 | tail -6 | tee $OUT/latency_${TAG}.log
