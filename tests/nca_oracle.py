"""TEST INFRASTRUCTURE.  The NCA training step (tensorfrost_b200/nca_dp.py: grad program -> apply program) on the REFERENCE's
C++/OpenMP backend (oracle/_ref, strict flags), for live parity runs on the GPU box and for tests/golden/nca_step_smooth.npz.

usage: python tests/nca_oracle.py <out.npz> <global_batch> <grid> <pool> <train_steps> <iters> <quantize 0|1>
Stores: losses[iters], flat0 (gradients + loss of the first iteration), ids, lr and the configuration."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def batch_ids(batch, pool):
    return np.random.default_rng(123).choice(pool, batch, replace=False).astype(np.int32)


LR = 0.002


def main():
    out = sys.argv[1]
    batch, grid, pool, steps, iters, quantize = (int(v) for v in sys.argv[2:8])
    os.environ.setdefault("OMP_NUM_THREADS", str(len(os.sched_getaffinity(0))))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import TensorFrost as tf
    tf.initialize(tf.cpu, "-O3 -fopenmp -include math.h")
    from tensorfrost_b200 import nca_dp
    tr = nca_dp.NcaTrainer(tf, global_batch=batch, grid=grid, pool_size=pool, train_steps=steps, quantize=bool(quantize))
    ids = batch_ids(batch, pool)
    losses, flat0 = [], None
    for it in range(iters):
        losses.append(tr.step(batch_ids=ids, lr=LR, read_loss=True))
        if it == 0:
            flat0 = np.array(tr.last_flat.numpy)
    np.savez_compressed(out, losses=np.array(losses, np.float64), flat0=flat0, ids=ids, lr=np.array(LR), global_batch=np.array(batch), grid=np.array(grid),
                        pool_size=np.array(pool), train_steps=np.array(steps), quantize=np.array(quantize))
    print("losses", losses)


if __name__ == "__main__":
    main()
