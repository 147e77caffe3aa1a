#!/bin/bash
# Round-2 final evidence pass (1 GPU): full gpu test suite, default bench (+ CPU reference arm), ncu launch list + conv
# traffic of one step, --set full captures of the dominant conv instances, batch-1 latency.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q > $OUT/z_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/z_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/z_bench_default.json 2> $OUT/z_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>> $OUT/z_bench.err | grep "^{" | tail -1 > $OUT/z_bench_reference.json
FULL=1 bash tools/ncu_conv.sh r2_final > $OUT/z_ncu.log 2>&1
python tools/layer_table.py kernels $OUT/launches_r2_final.csv > $OUT/z_kernels.md
python tools/layer_table.py convs $OUT/conv_traffic_r2_final.csv > $OUT/z_convs.md
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_tc -s 3 -c 1 -f -o $OUT/r2_stem_full \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/z_ncu_stem.log 2>&1
timeout 300 python tools/latency_bench.py > $OUT/z_latency.jsonl 2>&1
tail -3 $OUT/z_pytest.log; python -c "
import json
j=json.loads([l for l in open('$OUT/z_bench_default.json') if l.startswith('{')][-1]); print('default', round(j['value'],1), round(j['ms_per_step'],2), round(j['e2e']['value'],1), round(j['roofline']['frac'],4), j['roofline']['traffic'], j['clocks'], j['cpu_baseline']['value'])
j=json.loads(open('$OUT/z_bench_reference.json').read()); print('reference', round(j['value'],2), j['cpu_baseline']['cores'], j['cpu_baseline']['kind'])"
tail -4 $OUT/z_convs.md; cat $OUT/z_latency.jsonl | tail -3; ls -la $OUT/*.ncu-rep
