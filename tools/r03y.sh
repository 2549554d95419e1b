#!/bin/bash
# r03y: the round's last evidence run on the final library: sanitizers on the reworked gi_continue, ncu --set full of the GI kernels, the
# re-queued reflection kernel and the a-trous kernel, instruction counts for profiles/issue.json, whole GPU suite, bench line, smoke
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool racecheck python tools/sanitize_run.py > gpurun_out/r03y_sanitizer_racecheck.log 2>&1; tail -2 gpurun_out/r03y_sanitizer_racecheck.log
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_run.py > gpurun_out/r03y_sanitizer_memcheck.log 2>&1; tail -2 gpurun_out/r03y_sanitizer_memcheck.log
for k in gi_gen_trace0 gi_continue; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_r03y_$k python tools/gi_probe.py 3 > gpurun_out/r03y_ncu_$k.log 2>&1
done
for k in refl_gen_trace svgf_spatial_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/prof_r03y_$k python tools/denoise_probe.py 2 > gpurun_out/r03y_ncu_$k.log 2>&1
done
timeout 300 ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r03y_inst_gi.csv python tools/gi_probe.py 3 > gpurun_out/r03y_inst.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r03y_pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r03y_bench_n1_driver_flags.json 2> gpurun_out/r03y_bench_n1.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r03y_bench_n1_driver_flags.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['pass_ms'])"; tail -2 gpurun_out/r03y_bench_n1.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r03y_smoke.log
