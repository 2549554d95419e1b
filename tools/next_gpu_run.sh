# First GPU call of the next round: everything that was built after this round's GPU budget was spent.
#   gpurun --timeout 900 -- 'bash tools/next_gpu_run.sh r02a'
# 1. the GPU tests that have never run on a GPU (SVGF denoiser, shadow filters, frame-level calls, relief parallax, render_frame + material pass)
# 2. timings of every f1 / f2 pass at 1080p (library CUDA events)   3. ncu launch list + full-set captures of the new kernels
R=${1:-r02a}
mkdir -p gpurun_out
# (the GPU tests of these kernels ran green in the round-1 driver run: GPUTEST_r01.json)
timeout 120 python tools/material_probe.py 30 > gpurun_out/${R}_material_probe.json 2> gpurun_out/${R}_material_probe.err; cat gpurun_out/${R}_material_probe.json
timeout 120 python tools/df_probe.py 30 > gpurun_out/${R}_df_probe.json 2> gpurun_out/${R}_df_probe.err; cat gpurun_out/${R}_df_probe.json
timeout 180 python tools/denoise_probe.py 20 > gpurun_out/${R}_denoise_probe.json 2> gpurun_out/${R}_denoise_probe.err; cat gpurun_out/${R}_denoise_probe.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_denoise_$R.csv python tools/denoise_probe.py 2 > gpurun_out/ncu_denoise_$R.log 2>&1
for k in svgf_initial_kernel svgf_temporal_kernel svgf_variance_kernel svgf_spatial_kernel shadow_temporal_kernel shadow_filter_kernel; do
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_${R}_$k python tools/denoise_probe.py 2 >> gpurun_out/ncu_denoise_$R.log 2>&1
done
# the material pass, default kernel (launch 4 of the probe) and the opt-in quad-shuffle instantiation (the probe's second half: 2 + 3 launches in, +3 warm-up)
timeout 150 ncu --set full --clock-control none --import-source on -k regex:gbuffer_kernel -s 3 -c 1 -f -o gpurun_out/prof_${R}_gbuffer_default python tools/material_probe.py 2 >> gpurun_out/ncu_denoise_$R.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:gbuffer_kernel -s 8 -c 1 -f -o gpurun_out/prof_${R}_gbuffer_quad_shuffle python tools/material_probe.py 2 >> gpurun_out/ncu_denoise_$R.log 2>&1
ls gpurun_out | grep $R | tr '\n' ' '
