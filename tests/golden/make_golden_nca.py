"""Golden fixture for the NCA training step: the REFERENCE's C++/OpenMP backend (oracle/_ref, compiled without fast-math)
runs (a) the reference's own single-program step and (b) the split grad/apply step of tensorfrost_b200/nca_dp.py on a small
seeded configuration.  Stored: the loss sequences of both, and the flat [gradients..., loss] tensor of the first split step.

usage: python tests/golden/make_golden_nca.py        (needs oracle/_ref; run in the build container)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

CONFIG = dict(global_batch=4, grid=24, pool_size=16, train_steps=3)
IDS = np.array([3, 7, 1, 12], np.int32)
LR = 0.002
ITERS = 3


def run(tf, mono):
    from tensorfrost_b200 import nca_dp
    tr = nca_dp.NcaTrainer(tf, mono=mono, **CONFIG)
    losses, flat0, state0 = [], None, None
    for it in range(ITERS):
        losses.append(tr.step(batch_ids=IDS, lr=LR, read_loss=True))
        if it == 0 and not mono:
            flat0 = np.array(tr.last_flat.numpy)
            state0 = np.array(tr.last_state.numpy)
    return np.array(losses, np.float64), flat0, state0


def main():
    import TensorFrost as tf
    tf.initialize(tf.cpu, "-O3 -fopenmp -include math.h")
    mono_losses, _, _ = run(tf, True)
    split_losses, flat0, state0 = run(tf, False)
    np.savez_compressed(os.path.join(HERE, "nca_step.npz"), mono_losses=mono_losses, split_losses=split_losses, flat0=flat0, state0=state0,
                        ids=IDS, lr=np.array(LR), **{k: np.array(v) for k, v in CONFIG.items()})
    print("mono", mono_losses, "split", split_losses, "flat", flat0.shape, "state", state0.shape)


if __name__ == "__main__":
    main()
