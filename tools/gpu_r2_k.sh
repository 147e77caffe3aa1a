#!/bin/bash
# Round-2 GPU pass K (1 GPU): where the fused stem kernel's time goes (timing-only diagnostics, results are wrong by design).
set -u
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
for d in 0 1 2 4 3 7; do
  ACR_B200_STEM_DIAG=$d timeout 300 $B > $OUT/k_bench_$d.json 2>> $OUT/k_bench.err
  python -c "
import json
try:
    j=json.loads([l for l in open('$OUT/k_bench_$d.json') if l.startswith('{')][-1]); print('diag $d', round(j['value'],1), j['profile_ms_by_kind'].get('11'))
except Exception as e: print('diag $d', 'ERR', e); print(open('$OUT/k_bench.err').read()[-600:])
"
done
