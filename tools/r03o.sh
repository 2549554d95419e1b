#!/bin/bash
# r03o: denoiser after branch-free texel wraps, short exp / pow evaluations and integer face-normal weights
mkdir -p gpurun_out
python tools/denoise_probe.py 10 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k: round(v['ms'],4) for k,v in d['passes'].items()})" > gpurun_out/r03o_probe.txt 2>&1
cat gpurun_out/r03o_probe.txt
timeout 900 python -m pytest tests/test_svgf_denoise.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r03o_pytest.txt
