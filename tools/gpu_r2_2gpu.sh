#!/bin/bash
# Round-2 2-GPU pass: fused vertex all-gather (double-buffered slots, flags, counts) vs NCCL, 10 steps; bench A/B at N=2.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/g2_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_dist2.py -q -s > $OUT/g2_dist_test.log 2>&1; echo "dist test exit $?" | tee -a $OUT/g2_dist_test.log
for G in fused nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 10 --warmup 3 --gather $G > $OUT/g2_bench_$G.json 2> $OUT/g2_bench_$G.err
  echo "bench $G exit $?" | tee -a $OUT/g2_dist_test.log
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/g2_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/g2_smoke.log; tail -3 $OUT/g2_smoke.log
tail -5 $OUT/g2_dist_test.log
for G in fused nccl; do python -c "
import json
try:
    j=json.loads([l for l in open('$OUT/g2_bench_$G.json') if l.startswith('{')][-1]); print('$G', round(j['value'],1), round(j['ms_per_step'],2), j.get('gather_check'), j.get('ms_per_step_by_rank'), j['config']['parallelism'][:90])
except Exception as e: print('$G', 'ERR', e); print(open('$OUT/g2_bench_$G.err').read()[-1500:])
"; done
