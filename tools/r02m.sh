# r02m: (1) one rank's share of an 8-way frame on one GPU (--emulate 8): pipes and GI sort width; (2) full GPU suite; (3) bench N=1
# (aux passes after the GGX table), configs 4 and 5; (4) instruction counts for profiles/issue.json
mkdir -p gpurun_out
emu() { # tag env... -- args
  tag=$1; shift
  envs=""; while [ "$1" != "--" ]; do envs="$envs $1"; shift; done; shift
  env $envs timeout 200 python bench.py --emulate 8 --no-cpu-baseline --no-aux "$@" > gpurun_out/r02m_emu_$tag.json 2> gpurun_out/r02m_emu_$tag.err
  python - "$tag" <<'PY'
import json, sys
n='gpurun_out/r02m_emu_%s.json' % sys.argv[1]
try:
    d=json.loads([l for l in open(n) if l.startswith('{')][-1])
    print(sys.argv[1], 'ms/step', round(d['ms_per_step'],4), 'x8 value', round(d['value']*8) if False else round(d['value']), d['pass_ms']['primary'], d['pass_ms']['shadow'], d['pass_ms']['diffuse'])
except Exception as e: print(n, 'ERR', e)
PY
}
emu p8_s20 -- --pipes 8 --steps 20 --warmup 5
emu p8_s200 -- --pipes 8 --steps 200 --warmup 10
emu p8_sort4_s20 VXPT_GI_SORT=4 -- --pipes 8 --steps 20 --warmup 5
emu p8_sort4_s200 VXPT_GI_SORT=4 -- --pipes 8 --steps 200 --warmup 10
emu p6_s20 -- --pipes 6 --steps 20 --warmup 5
emu p10_s20 -- --pipes 10 --steps 20 --warmup 5
emu p4_s20 -- --pipes 4 --steps 20 --warmup 5
emu p4_s200 -- --pipes 4 --steps 200 --warmup 10
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02m_pytest.log 2>&1; tail -3 gpurun_out/r02m_pytest.log
timeout 200 ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02m_inst_gi.csv python tools/gi_probe.py 3 > /dev/null 2>&1
python tools/make_issue_json.py gpurun_out/r02m_inst_gi.csv; cp profiles/issue.json gpurun_out/r02m_issue.json
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02m_bench_n1.json 2> gpurun_out/r02m_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02m_bench_n1.json')); print(d['value'], d['ms_per_step'], d['pass_ms'], d['roofline']['frac'], d['roofline'].get('issue_frac'), {k:v.get('issue_frac') for k,v in d['roofline_all'].items()}, d['e2e']['value'], d['e2e']['pcie_gbs_per_gpu'])"; tail -3 gpurun_out/r02m_bench_n1.err
for c in 4 5; do
  timeout 400 python bench.py --config $c --steps 50 --warmup 5 > gpurun_out/r02m_bench_config$c.json 2> gpurun_out/r02m_bench_config$c.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/r02m_bench_config$c.json')); print('config $c', d['value'], d['ms_per_step'], d['pass_ms'], d['roofline']['frac'], d.get('rebuild'), d['e2e']['value'])"; tail -3 gpurun_out/r02m_bench_config$c.err
done
