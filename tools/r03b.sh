# r03b: end-to-end loop with the frame call recorded into CUDA graphs, N = 8 and 4 (driver flags)
mkdir -p gpurun_out
run() { # N steps warmup tag extra...
  N=$1; S=$2; Wm=$3; T=$4; shift 4
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29730+RANDOM%200))"
  timeout 300 $TR bench.py --gpus $N --steps $S --warmup $Wm "$@" > gpurun_out/r03b_bench_n${N}_$T.json 2> gpurun_out/r03b_bench_n${N}_$T.err
  python - "$N" "$T" <<'PY'
import json, sys
n='gpurun_out/r03b_bench_n%s_%s.json' % (sys.argv[1], sys.argv[2])
try:
    d=json.loads([l for l in open(n) if l.startswith('{')][-1])
    print(n, round(d['value']), round(d['ms_per_step'],4), 'gathered', d.get('gathered_ok'), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), round(d['e2e']['pcie_gbs_per_gpu'],1), d['e2e']['submit'][:30])
except Exception as e: print(n, 'ERR', e)
PY
  grep -iE "error|fallback|failed|Traceback" gpurun_out/r03b_bench_n${N}_$T.err | head -3
}
run 8 20 5 driver
run 4 20 5 driver
run 8 100 5 s100
