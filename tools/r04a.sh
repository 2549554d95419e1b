#!/bin/bash
# r04a: denoiser with per-kernel register caps (temporal / variance 64, shadow filters 48, a-trous and pre-pass uncapped) and literal exp constants
mkdir -p gpurun_out
for i in 1 2; do
  timeout 120 python tools/denoise_probe.py 10 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k: round(v['ms'],4) for k,v in d['passes'].items() if k not in ('material','reflection')})"
done | tee gpurun_out/r04a_denoise.txt
timeout 900 python -m pytest tests/test_svgf_denoise.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r04a_pytest.txt
