"""Run under torchrun on N >= 2 GPUs (not collected by pytest: `python -m torch.distributed.run --nproc-per-node N tests/dp_exchange_check.py`).
Checks the one-shot peer-memory allreduce (csrc/comm.cu) against the exact sum and against ncclAllReduce on the same data, over many
consecutive exchanges (both slot parities, back-to-back launches with no host synchronisation in between), and that every rank ends
with bit-identical results.  Prints one line per rank-0 check and EXCHANGE-OK at the end."""
import hashlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    import tensorfrost_b200
    from tensorfrost_b200 import nca_dp
    tf = tensorfrost_b200.load()
    method = nca_dp.init_comm(tf, rank, world)
    assert method == "peer", f"peer exchange not available: {method}"
    n = 7821
    rng = np.random.default_rng(1234)
    base = rng.standard_normal((world, 64, n)).astype(np.float32)  # every rank knows every rank's data: exact expected sums
    digest = hashlib.sha256()
    for it in range(64):
        mine = tf.cuda_tensor(base[rank, it])
        tf.cuda_allreduce(mine, 1.0 / world, "peer")
        if it % 8 == 7:  # most exchanges run back to back, some are read immediately
            got = tf.cuda_numpy(mine)
            acc = base[0, it].copy()
            for r in range(1, world):
                acc = acc + base[r, it]        # rank-order fp32 sum: what the kernel computes
            want = acc * np.float32(1.0 / world)
            assert np.array_equal(got, want), f"rank {rank} iteration {it}: peer allreduce differs from the rank-order sum"
            digest.update(got.tobytes())
            other = tf.cuda_tensor(base[rank, it])
            tf.cuda_allreduce(other, 1.0 / world, "nccl")
            np.testing.assert_allclose(tf.cuda_numpy(other), got, rtol=1e-6, atol=1e-6)
    tf.cuda_synchronize()
    # timing: 200 back-to-back exchanges of the NCA payload
    t = tf.cuda_tensor(base[rank, 0])
    for methodname in ("peer", "nccl"):
        for _ in range(10):
            tf.cuda_allreduce(t, 1.0, methodname)
        tf.cuda_synchronize()
        dist.barrier()
        tf.cuda_timer_begin()
        for _ in range(200):
            tf.cuda_allreduce(t, 1.0, methodname)
        ms = tf.cuda_timer_end()
        if rank == 0:
            print(f"[exchange] {methodname}: {ms / 200 * 1e3:.1f} us per allreduce of {n} floats at {world} ranks", flush=True)
    every = [None] * world
    dist.all_gather_object(every, digest.hexdigest())
    assert len(set(every)) == 1, "ranks hold different results"
    dist.barrier()
    if rank == 0:
        print("EXCHANGE-OK", world, flush=True)
    tf.cuda_comm_destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    import faulthandler
    faulthandler.dump_traceback_later(240, exit=True)
    main()
