"""Multi-GPU row-slab gather on real devices (needs >= 2 GPUs; the driver's single-GPU box skips it): tools/p2p_check.py under
torch.distributed.run — gathered planes == a full-frame render, for the peer-to-peer store path and for the NCCL all-gather."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("mode,planes", [("p2p", "texel"), ("p2p", "f32"), ("p2pcopy", "texel"), ("nccl", "texel")])
def test_gathered_planes_equal_a_full_frame_render(mode, planes):
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29741",
           os.path.join(ROOT, "tools", "p2p_check.py"), mode, planes]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and ": OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
