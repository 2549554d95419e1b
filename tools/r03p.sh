#!/bin/bash
# r03p: whole GPU suite and the bench line as the driver runs it, on the library with the denoiser instruction diet
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r03p_pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r03p_bench_n1_driver_flags.json 2> gpurun_out/r03p_bench_n1.err
tail -c 3000 gpurun_out/r03p_bench_n1_driver_flags.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r03p_smoke.log
