"""Exercise tfcuda_matmul modes 0/1/2 through the C-ABI on a set of shapes; print relative errors vs float64 numpy and timings.
Each case runs in this one process; run under `timeout`."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorfrost_b200 import abi

abi.init(-1)
lib = abi.lib()


def run(m, n, k, mode, seed=0, iters=0, dist="uniform"):
    rng = np.random.default_rng(seed)
    if dist == "uniform":
        a, b = rng.random((m, k), dtype=np.float32), rng.random((k, n), dtype=np.float32)
    else:
        a, b = rng.standard_normal((m, k)).astype(np.float32), rng.standard_normal((k, n)).astype(np.float32)
    da, db, dc = abi.DeviceArray(a), abi.DeviceArray(b), abi.DeviceArray(np.full((m, n), np.nan, np.float32))
    abi.check(lib.tfcuda_matmul(da.ptr, db.ptr, dc.ptr, 1, m, n, k, mode), "matmul")
    abi.check(lib.tfcuda_sync(), "sync")
    c = dc.get()
    ms = None
    if iters:
        lib.tfcuda_timer_begin()
        for _ in range(iters):
            lib.tfcuda_matmul(da.ptr, db.ptr, dc.ptr, 1, m, n, k, mode)
        t = abi.f32()
        abi.check(lib.tfcuda_timer_end(t), "timer")
        ms = t.value / iters
    if m * n * k <= 2 ** 31:
        want = a.astype(np.float64) @ b.astype(np.float64)
        err = float(np.max(np.abs(c - want)) / np.max(np.abs(want)))
        fro = float(np.linalg.norm(c - want) / np.linalg.norm(want))
    else:
        # spot-check 64 rows
        rows = rng.integers(0, m, 64)
        want = a[rows].astype(np.float64) @ b.astype(np.float64)
        err = float(np.max(np.abs(c[rows] - want)) / np.max(np.abs(want)))
        fro = float(np.linalg.norm(c[rows] - want) / np.linalg.norm(want))
    nan = int(np.isnan(c).sum())
    print(f"mode {mode} {m}x{n}x{k} {dist}: max-rel {err:.3e} fro-rel {fro:.3e} nan {nan}" + (f"  {ms:.3f} ms  {2.0 * m * n * k / ms / 1e9:.1f} TFLOP/s" if ms else ""), flush=True)
    return err


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    shapes = [(128, 32, 32), (128, 256, 32), (128, 256, 64), (256, 512, 128), (1000, 200, 48), (4096, 128, 48), (4096, 12, 128), (1024, 1024, 1024), (300, 260, 1000)]
    if which in ("all", "small"):
        for mode in (0, 1):
            for (m, n, k) in shapes:
                run(m, n, k, mode)
            run(512, 384, 256, mode, dist="normal")
    if which in ("all", "big"):
        for mode in (2, 0, 1):
            run(8192, 8192, 8192, mode, iters=5)
        run(128 * 128 * 64, 128, 48, 0, iters=5)
        run(128 * 128 * 64, 128, 48, 1, iters=5)
        run(128 * 128 * 64, 12, 128, 1, iters=5)
