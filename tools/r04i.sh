#!/bin/bash
# r04i: evidence on the final library of the round: whole GPU suite, smoke, the bench line under the driver's flags
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r04i_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r04i_smoke.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r04i_bench_n1_driver_flags.json 2> gpurun_out/r04i_bench_n1.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r04i_bench_n1_driver_flags.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['pass_ms'])"; tail -2 gpurun_out/r04i_bench_n1.err
