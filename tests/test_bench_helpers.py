"""bench.py host logic that needs no GPU: the reference arm (the reference's compiled shaders on the host cores, or the oracle port) prints the
contract's JSON line; slab sampling covers the frame; NUMA binding degrades gracefully."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env={**os.environ, "RANK": "0"})
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "Mrays/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["metric"].startswith("Mrays/s") and "workload" in line["config"]
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the other ranks of a torchrun launch exit without work
    other = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                           timeout=120, env={**os.environ, "RANK": "1"})
    assert other.returncode == 0 and other.stdout.strip() == ""


def test_slab_sampling_and_affinity_helpers():
    import bench

    class Probe(bench.CpuReference):
        def __init__(self, stride):
            self.stride = stride

    assert sum(re - rb for rb, re in Probe(1).slabs()) == bench.HEIGHT
    rows3 = Probe(3).slabs()
    assert rows3[0] == (0, 8) and rows3[1] == (24, 32) and sum(re - rb for rb, re in rows3) == 8 * 45
    before = os.sched_getaffinity(0)
    bench.bind_near_gpu(0)                      # no NVML / no GPU here: must not raise, must not shrink the CPU set to nothing
    assert len(os.sched_getaffinity(0)) >= 1
    os.sched_setaffinity(0, before)
