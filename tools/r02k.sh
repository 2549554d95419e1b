# r02k: min-plus step latency probe (DPX vs integer vs half2); vxpt_mg_* on one GPU
mkdir -p gpurun_out
./tools/probes/dpx_latency > gpurun_out/r02k_dpx_latency.txt 2>&1; cat gpurun_out/r02k_dpx_latency.txt
timeout 600 python -m pytest tests/test_mg_frame.py tests/test_z_material_extras.py -m gpu -x -q > gpurun_out/r02k_pytest.log 2>&1; tail -3 gpurun_out/r02k_pytest.log
