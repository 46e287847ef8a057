"""Zero-copy interop of device tensors: tf.TensorMemory exposes __cuda_array_interface__ (v3) and DLPack, so torch /
cupy / numba see the backend's buffers without a copy; tf.cuda_from_device_array imports any such array with one device-to-device
copy; and the reference's own tf.tensor(np) / .numpy take the bulk path for contiguous 4-byte arrays (PyTensorMemory.cpp:15-84
otherwise converts element by element).  Runs in a subprocess: torch has to be imported before TensorFrost in this image."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
pytestmark = pytest.mark.gpu

SCRIPT = r'''
import sys, time
sys.path.insert(0, %r)
import numpy as np
import torch
import tensorfrost_b200
tf = tensorfrost_b200.load()
a = np.arange(6 * 1000, dtype=np.float32).reshape(6, 1000)
t = tf.tensor(a)                                   # reference API, bulk path underneath
cai = t.__cuda_array_interface__
assert cai["shape"] == (6, 1000) and cai["typestr"] == "<f4" and cai["version"] == 3 and cai["data"][0] == tf.cuda_device_ptr(t)
view = torch.as_tensor(t, device="cuda")           # zero-copy through __cuda_array_interface__
assert view.data_ptr() == tf.cuda_device_ptr(t) and tuple(view.shape) == (6, 1000)
view.mul_(2.0)
torch.cuda.synchronize()
assert np.array_equal(np.array(t.numpy), a * 2)    # the backend sees torch's write: same memory
d = torch.from_dlpack(t)                           # zero-copy through DLPack
assert d.data_ptr() == tf.cuda_device_ptr(t) and d.dtype == torch.float32 and d.device.type == "cuda"
d.add_(1.0)
torch.cuda.synchronize()
assert np.array_equal(tf.cuda_numpy(t), a * 2 + 1)
del view, d
for dtype, tdtype in ((np.int32, torch.int32),):
    x = tf.tensor(np.arange(10, dtype=dtype))
    assert torch.from_dlpack(x).dtype == tdtype
src = torch.arange(5000, dtype=torch.float32, device="cuda").reshape(50, 100) * 0.5
torch.cuda.synchronize()
imported = tf.cuda_from_device_array(src)          # device-to-device import
assert imported.shape == [50, 100] or tuple(imported.shape) == (50, 100)
assert np.array_equal(np.array(imported.numpy), src.cpu().numpy())
# bulk path of the reference API: 2^24 elements through tf.tensor / .numpy in well under a second each (per-element: ~10 s)
big = np.random.default_rng(0).random(1 << 24, dtype=np.float32)
t0 = time.perf_counter(); tb = tf.tensor(big); up = time.perf_counter() - t0
t0 = time.perf_counter(); back = np.array(tb.numpy); down = time.perf_counter() - t0
assert np.array_equal(back, big)
assert up < 1.0 and down < 1.0, (up, down)
strided = np.arange(40, dtype=np.float32).reshape(5, 8)[:, ::2]     # non-contiguous: the reference path, same values
assert np.array_equal(np.array(tf.tensor(strided).numpy), strided)
assert np.array_equal(np.array(tf.tensor(np.arange(7, dtype=np.float64)).numpy), np.arange(7, dtype=np.float32))
print("INTEROP-OK", up, down)
''' % ROOT


def test_cuda_array_interface_dlpack_and_bulk_paths(tmp_path):
    script = tmp_path / "interop.py"
    script.write_text(SCRIPT)
    r = subprocess.run([sys.executable, str(script)], cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "INTEROP-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
