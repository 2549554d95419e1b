# r03d: compute-sanitizer on the final kernels (distance-field early-outs, parallel carries, reflection wavefront, vxpt_mg)
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck python tools/sanitize_run.py > gpurun_out/r03d_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/r03d_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck python tools/sanitize_run.py > gpurun_out/r03d_sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/r03d_sanitizer_racecheck.log
timeout 1200 compute-sanitizer --tool synccheck python tools/sanitize_run.py > gpurun_out/r03d_sanitizer_synccheck.log 2>&1; tail -3 gpurun_out/r03d_sanitizer_synccheck.log
