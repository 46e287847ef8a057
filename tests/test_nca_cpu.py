"""CPU tests of the data-parallel NCA trainer (tensorfrost_b200/nca_dp.py): host-side sharding / packing logic, and the
whole split step (grad program -> exchange -> apply program) at world_size 2 over gloo, traced and executed on the
reference's own C++/OpenMP backend (oracle/_ref) so no GPU is needed.  The exchange there is a gloo allreduce; on the
GPU it is NCCL (tests/test_nca_gpu.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from tensorfrost_b200 import nca_dp  # noqa: E402

HAVE_ORACLE = os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "TensorFrost"))
HAVE_WORKLOAD = os.path.exists(os.path.join(ROOT, "build", "workloads", "nca_program.py.txt"))


def test_shard_sizes():
    assert nca_dp.shard_sizes(256, 1024, 1) == (256, 1024)
    assert nca_dp.shard_sizes(256, 1024, 8) == (32, 128)
    with pytest.raises(ValueError):
        nca_dp.shard_sizes(256, 1024, 3)
    with pytest.raises(ValueError):
        nca_dp.shard_sizes(64, 32, 1)


def test_batch_ids_unique_and_in_shard():
    rng = np.random.default_rng(0)
    for _ in range(20):
        ids = nca_dp.draw_batch_ids(rng, 128, 32)
        assert ids.dtype == np.int32 and len(set(ids.tolist())) == 32 and ids.min() >= 0 and ids.max() < 128


def test_flat_layout_roundtrip():
    shapes = [(48, 128), (128,), (128, 12), (12,)]
    offsets, total = nca_dp.flat_layout(shapes)
    assert offsets == [0, 6144, 6272, 7808] and total == 7821  # 7,820 gradient floats + the loss
    rng = np.random.default_rng(1)
    arrays = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    flat = nca_dp.pack_flat(arrays, 0.25)
    back, loss = nca_dp.unpack_flat(flat, shapes)
    assert loss == 0.25 and all(np.array_equal(a, b) for a, b in zip(arrays, back))


def test_lr_schedule_matches_train_py():
    assert nca_dp.lr_schedule(0) == 0.05
    assert abs(nca_dp.lr_schedule(500) - 0.035) < 1e-12
    assert abs(nca_dp.lr_schedule(2500) - 0.006) < 1e-12
    assert nca_dp.lr_schedule(10 ** 6) == 0.002


WORKER = r'''
import os, sys
import numpy as np
import torch                      # before TensorFrost (the image segfaults the other way round)
import torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "oracle", "_ref"))
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
import TensorFrost as tf
tf.initialize(tf.cpu, "-O3 -fopenmp -include math.h")
from tensorfrost_b200 import nca_dp

def gloo_exchange(flat, world):
    t = torch.from_numpy(np.array(flat.numpy, dtype=np.float32))
    dist.all_reduce(t)
    return tf.tensor((t / world).numpy())

mode = {mode!r}
# The ORACLE compiles every program from the fixed path /tmp/generated_lib_<id>.cpp (Backends/CPU/KernelCompiler.cpp:93-113): two ranks
# tracing at the same instant overwrite each other's source (the race the CUDA backend's host-program cache removes, DESIGN.md 12).
# On the unmodified reference the ranks therefore take turns: tf.compile runs g++ inside the constructor, before any collective.
import fcntl
lock = open(os.path.join(os.path.dirname({out!r}), "compile.lock"), "w")
fcntl.flock(lock, fcntl.LOCK_EX)
# "same": both ranks see the SAME data and RNG stream, so the averaged gradient equals the single-rank one.
tr = nca_dp.NcaTrainer(tf, global_batch=4 * world if mode == "same" else 4, grid=24, pool_size=16 * world if mode == "same" else 16,
                       train_steps=2, rank=rank, world=world, exchange=gloo_exchange, rank_seed_offset=0 if mode == "same" else 1000003)
fcntl.flock(lock, fcntl.LOCK_UN)
ids = np.array([3, 7, 1, 12], np.int32)[: tr.batch]
losses = [tr.step(batch_ids=ids if mode == "same" else None, lr=0.002, read_loss=True) for _ in range(2)]
params = tr.parameters_numpy()
np.savez({out!r} + f".{{rank}}.npz", losses=np.array(losses), batch=np.array(tr.batch), pool=np.array(tr.pool_shard), *params)
dist.destroy_process_group()
'''


def _launch(tmp_path, mode, world):
    script = tmp_path / f"worker_{mode}.py"
    out = str(tmp_path / f"out_{mode}_{world}")
    script.write_text(WORKER.format(root=ROOT, mode=mode, out=out))
    port = 29500 + (os.getpid() % 2000)
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = [p.communicate(timeout=900)[0] for p in procs]
    for p, log in zip(procs, logs):
        assert p.returncode == 0, log[-3000:]
    results = []
    for r in range(world):
        with np.load(out + f".{r}.npz") as z:
            results.append({k: z[k] for k in z.files})
    return results


@pytest.mark.skipif(not (HAVE_ORACLE and HAVE_WORKLOAD), reason="needs oracle/_ref and the extracted NCA program (build())")
def test_dp_step_world2_gloo_same_data_equals_single_rank(tmp_path):
    two = _launch(tmp_path, "same", 2)
    one = _launch(tmp_path, "same", 1)
    # replicated, deterministic apply step: ranks hold bit-identical parameters with no broadcast
    for k in two[0]:
        if k.startswith("arr_"):
            assert np.array_equal(two[0][k], two[1][k]), f"{k} differs between ranks"
    # mean of two identical gradients == the gradient: same losses as the single-rank run
    np.testing.assert_allclose(two[0]["losses"], one[0]["losses"], rtol=1e-4)
    assert int(two[0]["batch"]) == 4 and int(two[0]["pool"]) == 16


@pytest.mark.skipif(not (HAVE_ORACLE and HAVE_WORKLOAD), reason="needs oracle/_ref and the extracted NCA program (build())")
def test_dp_step_world2_gloo_sharded_batch(tmp_path):
    two = _launch(tmp_path, "shard", 2)
    assert int(two[0]["batch"]) == 2 and int(two[0]["pool"]) == 8  # global batch 4 / pool 16 split over 2 ranks
    for k in two[0]:
        if k.startswith("arr_") and two[0][k].dtype == np.float32:  # weights, Adam moments, t; the uint32 RNG seed is per rank by design
            assert np.array_equal(two[0][k], two[1][k]), f"{k} differs between ranks"
    seeds = [[v for k, v in r.items() if k.startswith("arr_") and v.dtype == np.uint32] for r in two]
    assert not np.array_equal(seeds[0][0], seeds[1][0])
    # the exchanged loss slot is the mean over ranks: both ranks report the same number
    assert np.array_equal(two[0]["losses"], two[1]["losses"])
    assert np.all(np.isfinite(two[0]["losses"]))
