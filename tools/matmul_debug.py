import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensorfrost_b200 import abi
abi.init(-1)
lib = abi.lib()
np.set_printoptions(linewidth=200, precision=4, suppress=True)
def go(a, b, mode=0, tag=""):
    m, k = a.shape; n = b.shape[1]
    da, db, dc = abi.DeviceArray(a), abi.DeviceArray(b), abi.DeviceArray(np.full((m, n), np.nan, np.float32))
    rc = lib.tfcuda_matmul(da.ptr, db.ptr, dc.ptr, 1, m, n, k, mode)
    print(tag, "rc", rc, lib.tfcuda_last_error() if rc else "", "sync", lib.tfcuda_sync())
    c = dc.get()
    want = a.astype(np.float64) @ b.astype(np.float64)
    print(tag, "nonzero", int(np.count_nonzero(c)), "of", c.size, "max", float(np.abs(c).max()), "want max", float(np.abs(want).max()))
    print(c[:4, :8]); print(want[:4, :8])
    return c
m, n, k = 128, 32, 32
go(np.ones((m, k), np.float32), np.ones((k, n), np.float32), tag="ones")
a = np.zeros((m, k), np.float32); a[np.arange(m), np.arange(m) % k] = 1.0
b = np.arange(k * n, dtype=np.float32).reshape(k, n)
go(a, b, tag="select-rows")
a = np.arange(m * k, dtype=np.float32).reshape(m, k) % 7
b = np.zeros((k, n), np.float32); b[np.arange(n) % k, np.arange(n)] = 1.0
go(a, b, tag="select-cols")
