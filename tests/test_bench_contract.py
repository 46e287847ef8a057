"""CPU tests of bench.py's output contract: the CUDA arm's JSON line is assembled from a scripted stand-in for the device module (no
kernels run here, the numbers are meaningless: only the keys, types and bookkeeping are checked), and the reference arm is run for
real on the host cores at a small size."""
import json
import os
import subprocess
import sys
import types

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)


class _T:
    pass


class _ScriptedDevice:
    """Answers the tf.cuda_* calls bench_fluid makes; counts the host<->device copies of the e2e leg."""

    def __init__(self):
        self.uploads = self.downloads = 0

    def cuda_tensor(self, a): return _T()
    def cuda_synchronize(self): pass
    def cuda_launch_count(self): return 0
    def cuda_timer_begin(self): pass
    def cuda_timer_end(self): return 1.0
    def cuda_profile_reset(self): pass
    def cuda_profile_enable(self, on): pass
    def cuda_pinned_array(self, shape, dtype): return np.zeros(shape, np.float32)

    def cuda_profile_records(self):
        return [{"name": "kernel_0", "launches": 2, "total_ms": 0.26, "bytes": 2.0e8}, {"name": "kernel_14", "launches": 2, "total_ms": 0.13, "bytes": 4.0e8}]

    def cuda_upload(self, t, a): self.uploads += 1
    def cuda_download(self, t, a): self.downloads += 1
    def cuda_upload_async(self, t, a): self.uploads += 1
    def cuda_download_async(self, t, a): self.downloads += 1
    def cuda_wait_uploads(self): self.waits = getattr(self, "waits", 0) + 1
    def cuda_copy_sync(self): self.copy_syncs = getattr(self, "copy_syncs", 0) + 1
    def cuda_graph_stats(self): return {"enabled": True, "replays": 2, "exact_hits": 1, "patched": 0, "instantiated": 1, "eager_launches": 0}


def test_cuda_arm_line_has_every_contract_key(monkeypatch):
    import bench
    from tensorfrost_b200 import workloads
    monkeypatch.setattr(workloads, "load_fluid", lambda tf, n, m: (lambda *s: [_T()] * 7))
    dev = _ScriptedDevice()
    args = types.SimpleNamespace(size=2048, warmup=1, steps=2, no_verify=True)
    res = bench.bench_fluid(dev, None, 0, 1, args, bench.read_peaks())
    line = bench.make_line(args, 1, res)
    json.dumps(line)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "gpu_launches", "clocks", "roofline", "e2e"):
        assert key in line, key
    assert line["warmup"] >= 3 and line["steps"] == 2 and line["n_gpus"] == 1 and line["vs_baseline"] is None and line["scaling"] == "weak"
    assert "workload" in line["config"] and "model" not in line["config"]
    roof = line["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in roof, key
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12
    assert roof["kernel"] == "kernel_0" and roof["traffic"] == pytest.approx(70649088.0)  # profiles/r02_fluid_ncu_full_summary.csv: 50.36 + 20.29 MB
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] == 4 * 2048 * 2048 * 4 == e2e["d2h_bytes_per_step"] and e2e["unit"] == "GB/s"
    # every step uploads its 4 input fields and downloads its 4 result fields inside the timed region (copy streams); the same loop
    # runs the reported number of untimed steps first (at least bench.FLUID_WARMUP: graph instantiation must lie outside the timed region)
    total = args.steps + bench.fluid_warmup(args)
    assert line["warmup"] == bench.fluid_warmup(args) >= 8
    assert dev.uploads == 4 * total and dev.downloads == 4 * total and dev.waits == total and dev.copy_syncs == 2
    assert line["config"] == bench.fluid_config(2048)  # identical to the reference arm's (the driver compares the two arms' configs)
    assert bench.fluid_step_bytes(2048) == 816840772  # = the live count of the runtime profiler on the B200 (BENCH_r01.json)


def test_reference_arm_prints_one_json_line(tmp_path):
    if not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "TensorFrost")):
        pytest.skip("oracle/_ref (the reference module) is not built here")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--size", "256"],
                       cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "fused-kernel HBM GB/s" and line["unit"] == "GB/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    import bench
    assert line["config"] == bench.fluid_config(256)  # the same dict the CUDA arm prints (the driver compares the arms' configs)
    assert line["e2e"] == {"value": line["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
