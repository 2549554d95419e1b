#!/bin/bash
# r04j: four GPUs under the driver's flags on the final library, four frame pipes per GPU against the default eight (r04h: one rank's share of a
# sharded frame on one GPU runs 2-3.5 % faster over 20 steps with four)
mkdir -p gpurun_out
for p in 4 8; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29780 + p))"
  timeout 110 $TR bench.py --gpus 4 --steps 20 --warmup 5 --pipes $p > gpurun_out/r04j_bench_n4_pipes$p.json 2> gpurun_out/r04j_bench_n4_pipes$p.err
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/r04j_bench_n4_pipes$p.json') if l.startswith('{')][-1]); print('pipes $p', round(d['value']), d['ms_per_step'], d['config']['submit'][:30], d.get('gathered_ok'), round(d['e2e']['value']))"
  grep -iE "error|fallback|failed|Traceback" gpurun_out/r04j_bench_n4_pipes$p.err | head -3
done 2>&1 | tee gpurun_out/r04j_summary.txt
