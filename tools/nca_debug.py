"""Debug aid: per-gradient error of the split NCA step against the golden fixture under the current TFCUDA_* environment."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tensorfrost_b200
from tensorfrost_b200 import nca_dp
tf = tensorfrost_b200.load()
if os.environ.get("NCA_DEBUG_DIRTY"):
    # emulate the pytest context: earlier library tests left freed device memory full of positive floats
    from tensorfrost_b200 import abi
    abi.init(-1)
    rng = np.random.default_rng(5)
    keep = [abi.DeviceArray(rng.random(n, dtype=np.float32) + 0.5) for n in (1, 16, 2048, 3000, 60000, 300000, 1 << 20, 1 << 22, 1 << 24) for _ in range(3)]
    for d in keep:
        d.free()
    tf.cuda_synchronize()
g = np.load(os.path.join(ROOT, "tests", "golden", "nca_step.npz"))
saved = os.dup(1); os.dup2(os.open(os.devnull, os.O_WRONLY), 1)
tr = nca_dp.NcaTrainer(tf, mono=False, global_batch=int(g["global_batch"]), grid=int(g["grid"]), pool_size=int(g["pool_size"]), train_steps=int(g["train_steps"]))
loss = tr.step(batch_ids=g["ids"], lr=float(g["lr"]), read_loss=True)
flat = np.array(tr.last_flat.numpy)
os.dup2(saved, 1)
offsets, total = nca_dp.flat_layout(tr.grad_shapes)
env = {k: v for k, v in os.environ.items() if k.startswith("TFCUDA")}
out = [f"env={env} loss={loss:.6f} want={g['split_losses'][0]:.6f}"]
for off, shape in zip(offsets, tr.grad_shapes):
    n = int(np.prod(shape))
    a, b = flat[off:off + n].astype(np.float64), g["flat0"][off:off + n].astype(np.float64)
    out.append(f"  {shape}: max {np.abs(a-b).max()/max(np.abs(b).max(),1e-30):.2e} l2 {np.linalg.norm(a-b)/max(np.linalg.norm(b),1e-30):.2e}")
print("\n".join(out), flush=True)
