"""Times the library n-body step (current TFCUDA_NBODY_VARIANT) at 262144 bodies and checks it against the scalar variant's result."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tensorfrost_b200
tf = tensorfrost_b200.load()
rng = np.random.default_rng(0)
nb = 262144
x = tf.cuda_tensor((5.0 * rng.standard_normal((nb, 3))).astype(np.float32))
v = tf.cuda_tensor(np.zeros((nb, 3), np.float32))
for _ in range(2):
    out = tf.cuda_nbody_step(x, v)
tf.cuda_synchronize()
tf.cuda_timer_begin()
for _ in range(5):
    out = tf.cuda_nbody_step(x, v)
ms = tf.cuda_timer_end() / 5
vn = tf.cuda_numpy(out[1])
print(f"variant={os.environ.get('TFCUDA_NBODY_VARIANT','default')} ms={ms:.3f} Ginteractions/s={nb*nb/ms/1e6:.1f} checksum={np.abs(vn).sum():.6e} v0={vn[0]}", flush=True)
