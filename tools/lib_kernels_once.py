"""Runs every hand-written library kernel a few times at its BASELINE size, for `ncu` captures:
   ncu --set full --clock-control none --import-source on -k 'regex:onesweep|digit_histogram|gemm_tf32|reduce_rows|scan|nbody|matmul_tn' \
       --launch-skip 0 -c 40 -f -o gpurun_out/lib_full python tools/lib_kernels_once.py
Numbers printed by a run under ncu are not benchmark values."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tensorfrost_b200  # noqa: E402

tf = tensorfrost_b200.load()
rng = np.random.default_rng(0)
quick = "--quick" in sys.argv
medium = "--medium" in sys.argv  # ncu --set full saves and restores device memory around every replay pass: keep the footprint moderate
n = 1 << (22 if quick else (26 if medium else 28))
keys = tf.cuda_tensor(rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32))
vals = tf.cuda_tensor(np.arange(n, dtype=np.uint32))
for _ in range(2):
    tf.cuda_radix_sort(keys)
    tf.cuda_radix_sort(keys, vals)
del keys, vals
m = 2048 if quick else (4096 if medium else 8192)
a = tf.cuda_tensor(rng.random((m, m), dtype=np.float32))
b = tf.cuda_tensor(rng.random((m, m), dtype=np.float32))
for _ in range(2):
    tf.cuda_reduce(a, -1, "sum")
    tf.cuda_reduce(a, 0, "sum")
    tf.cuda_prefix_sum(a, -1)
    tf.cuda_matmul(a, b, 0)
    tf.cuda_matmul(a, b, 1)
nb = 32768 if quick else (65536 if medium else 262144)
x = tf.cuda_tensor((5.0 * rng.standard_normal((nb, 3))).astype(np.float32))
v = tf.cuda_tensor(np.zeros((nb, 3), np.float32))
tf.cuda_nbody_step(x, v)
ns = 1 << (20 if quick else 24)
dst = tf.cuda_tensor(np.zeros(1 << 20, np.float32))
idx = tf.cuda_tensor(rng.integers(0, 1 << 20, ns, dtype=np.int64).astype(np.int32))
src = tf.cuda_tensor(rng.random(ns, dtype=np.float32))
tf.cuda_scatter_add(dst, idx, src)
tf.cuda_synchronize()
print("ran", tf.cuda_launch_count(), "launches")
