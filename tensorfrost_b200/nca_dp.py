"""Data-parallel training of the reference's neural-cellular-automata example (BASELINE.json configs[4]) on 1/2/4/8 B200s.

The reference trains NCA with ONE compiled program per iteration (examples/ML/NCA/nca.py:213-231 `optimization_step`): forward
through `train_steps` CA steps, tf.grad, gradient-norm clipping and the Adam update all happen inside the traced program
(nca.py:124-168, Python/TensorFrost/optimizers.py:104-147), and train.py:142-144 feeds the returned parameters back.  It has no
distributed code.  Data parallelism needs the gradient exchange to sit BETWEEN tf.grad and the clip/Adam update, so the step is
split into two programs traced from the reference's own code:

  grad program    = CATrain.train_step with ModuleOptimizer._step replaced by "take tf.grad of every trainable parameter and
                    pack [grads..., loss] into one flat fp32 tensor" (the pattern of the reference's tests/autograd_test.py:9-23)
  exchange        = tf.cuda_allreduce(flat, 1/world): ncclAllReduce(sum) on the runtime stream + scale (libtfcuda comm.cu);
                    7,820 gradient floats + 1 loss for CAModel = 31 KB -> one latency-bound collective, no bucketing
  apply program   = ModuleOptimizer._step (unchanged reference code) with tf.grad replaced by "unpack from the flat tensor":
                    clip by norm, Adam, parameter update; replicated and deterministic, so parameters stay bit-identical across
                    ranks without any broadcast after step 0.

Sharding: the batch dimension.  Global batch B -> B/world samples per rank; every rank owns a pool shard of
POOL_SIZE/world states and draws its own batch ids; the model's RNG seed tensor is offset by rank so ranks do not draw the same
noise.  The "restart the worst sample" rule (nca.py:143-149) is applied per rank.  One process per GPU (the backend is a
process-global singleton), launched by torchrun; the NCCL id is broadcast through torch.distributed's store.

With world == 1 the split step computes exactly what the reference's single program computes (tests/test_nca_gpu.py pins it
against the oracle), and `mono=True` runs the reference's own unsplit program for comparison.
"""
import contextlib
import json
import os
import sys
import time

import numpy as np

from . import REPO_ROOT, workloads


# ----------------------------------------------------------------------------------------------------------------------
# pure host-side helpers (covered by the CPU / gloo tests)
# ----------------------------------------------------------------------------------------------------------------------
def shard_sizes(global_batch, pool_size, world):
    """Per-rank batch and pool shard.  Raises when the configuration does not divide (no silent truncation)."""
    if global_batch % world or pool_size % world:
        raise ValueError(f"global batch {global_batch} and pool {pool_size} must be divisible by world size {world}")
    if global_batch // world > pool_size // world:
        raise ValueError("per-rank batch exceeds the per-rank pool shard")
    return global_batch // world, pool_size // world


def draw_batch_ids(rng, pool_shard, batch):
    """train.py:141: np.random.choice(POOL_SIZE, BATCH_SIZE, replace=False), per rank on its own shard."""
    return rng.choice(pool_shard, batch, replace=False).astype(np.int32)


def flat_layout(shapes):
    """Offsets of each gradient inside the flat exchange buffer; the last slot is the loss."""
    offsets, total = [], 0
    for s in shapes:
        offsets.append(total)
        total += int(np.prod(s))
    return offsets, total + 1


def pack_flat(arrays, loss):
    """numpy restatement of the in-program packing (tests compare the two)."""
    return np.concatenate([np.asarray(a, np.float32).reshape(-1) for a in arrays] + [np.asarray([loss], np.float32)])


def unpack_flat(flat, shapes):
    offsets, total = flat_layout(shapes)
    assert flat.size == total
    return [flat[o:o + int(np.prod(s))].reshape(s) for o, s in zip(offsets, shapes)], float(flat[-1])


def lr_schedule(iteration, lrs=(0.05, 0.02, 0.01, 0.002), steps=(0, 1000, 2000, 3000)):
    """train.py:96-105 piecewise-linear schedule (train.py:142 multiplies it by 0.1)."""
    for i in range(len(steps) - 1):
        if steps[i] <= iteration < steps[i + 1]:
            t = (iteration - steps[i]) / (steps[i + 1] - steps[i])
            return lrs[i] * (1.0 - t) + lrs[i + 1] * t
    return lrs[-1]


# ----------------------------------------------------------------------------------------------------------------------
# tracing the two programs from the reference's code
# ----------------------------------------------------------------------------------------------------------------------
@contextlib.contextmanager
def _patched(obj, name, value):
    old = getattr(obj, name)
    setattr(obj, name, value)
    try:
        yield
    finally:
        setattr(obj, name, old)


def trainable(tf_module):
    ps, req = tf_module.parameters(), tf_module.requires_grads_list()
    return [p for p, r in zip(ps, req) if r]


def build_programs(tf, nca, train_steps):
    """Returns (grad_program, apply_program, mono_program_factory, grad_shapes)."""
    grad_shapes = []

    def make_train():
        return nca.CATrain(train_steps=train_steps)

    probe = make_train().opt.net
    for p in trainable(probe):
        grad_shapes.append(tuple(int(s) for s in p.shape))
    offsets, flat_size = flat_layout(grad_shapes)

    def grad_step():
        train = make_train()
        train.initialize_input()
        batch_ids = tf.input([nca.BATCH_SIZE], tf.int32)
        params = tf.input([-1], tf.float32)
        train.opt.net.fire_rate = params[0]
        flat = tf.buffer([flat_size], tf.float32)

        def take_gradients(opt, loss):
            for off, shape, p in zip(offsets, grad_shapes, trainable(opt.net)):
                g = tf.grad(loss, p)
                n = int(np.prod(shape))
                g = tf.reshape(g, [n])
                i, = g.indices
                flat[i + off] = g
            flat[flat_size - 1] = loss

        with _patched(tf.optimizers.ModuleOptimizer, "_step", take_gradients):
            loss, state = train.train_step(batch_ids)
        return [train.pool, train.opt.net.seed, flat, state]

    def apply_step():
        opt = make_train().opt
        opt.initialize_input()
        flat = tf.input([flat_size], tf.float32)
        params = tf.input([-1], tf.float32)
        opt.learning_rate = params[0]
        queue = []
        for off, shape in zip(offsets, grad_shapes):
            idx = tf.indices(list(shape))
            lin = idx[0]
            for d in range(1, len(shape)):
                lin = lin * shape[d] + idx[d]
            queue.append(flat[lin + off])

        def exchanged_gradient(loss, param):
            return queue.pop(0)

        # optimizers.py binds `tf` to the extension module: patch tf.grad there only, and only while tracing
        with _patched(tf.optimizers.tf, "grad", exchanged_gradient):
            opt._step(None)
        assert not queue, "the optimizer asked for fewer gradients than were exchanged"
        return opt.parameters()

    def mono_step():
        return nca.optimization_step()

    return grad_step, apply_step, mono_step, grad_shapes


class NcaTrainer:
    """One rank of the data-parallel NCA trainer (world == 1: plain single-GPU training)."""

    def __init__(self, tf, global_batch=256, grid=128, pool_size=1024, train_steps=25, rank=0, world=1, seed=0, mono=False,
                 channel_n=12, exchange=None, rank_seed_offset=1000003, quantize=True):
        """exchange(flat_tensor, world) -> flat_tensor: the gradient exchange; default = tf.cuda_allreduce (NCCL, in place).
        Tests on the CPU oracle backend inject a gloo exchange."""
        self.tf, self.rank, self.world, self.mono = tf, rank, world, mono
        self.exchange = exchange
        self.batch, self.pool_shard = shard_sizes(global_batch, pool_size, world)
        self.grid, self.train_steps = grid, train_steps
        nca = workloads.load_nca(tf, self.batch, grid, pool_size=self.pool_shard, train_steps=train_steps, channel_n=channel_n, quantize=quantize)
        self.nca = nca
        grad_step, apply_step, mono_step, self.grad_shapes = build_programs(tf, nca, train_steps)
        if mono:
            import functools
            nca.CATrain = functools.partial(nca.CATrain, train_steps=train_steps)
            self.mono_program = tf.compile(mono_step)
        else:
            self.grad_program = tf.compile(grad_step)
            self.apply_program = tf.compile(apply_step)
        self.trainer = nca.CATrain(train_steps=train_steps) if not mono else nca.CATrain()
        self.trainer.initialize_parameters()
        self.opt = self.trainer.opt
        self.model = self.opt.net
        # deterministic initial weights identical on every rank (tf.Parameter's own init is unseeded): train.py:60-80 otherwise
        rng = np.random.default_rng(seed)
        upload = getattr(tf, "cuda_tensor", tf.tensor)
        self._upload = upload
        hidden = int(self.model.fc1.shape[1])
        self.model.fc1 = upload((rng.standard_normal((channel_n * 4, hidden)) * np.sqrt(2.0 / (channel_n * 4))).astype(np.float32))
        self.model.fc1_bias = upload(np.zeros(hidden, np.float32))
        self.model.fc2 = upload(np.zeros((hidden, channel_n), np.float32))
        self.model.fc2_bias = upload(np.zeros(channel_n, np.float32))
        self.model.filters = upload(workloads.nca_filters())
        self.model.seed = upload(np.array([rank_seed_offset * rank], np.uint32))  # per-rank RNG stream
        self.trainer.image = upload(workloads.nca_target(grid, seed))
        self.trainer.pool = upload(workloads.nca_pool(self.pool_shard, grid, channel_n))
        self.ids_rng = np.random.default_rng(seed * 7919 + rank)
        self.iteration = 0
        self.fire_rate = float(nca.CELL_FIRE_RATE)

    # -- one training iteration -----------------------------------------------------------------------------------
    def step(self, batch_ids=None, lr=None, read_loss=False):
        tf = self.tf
        if batch_ids is None:
            batch_ids = draw_batch_ids(self.ids_rng, self.pool_shard, self.batch)
        if lr is None:
            lr = 0.1 * lr_schedule(self.iteration)
        self.iteration += 1
        if self.mono:
            outs = self.mono_program(self.trainer, batch_ids, np.array([lr, 0.0, self.fire_rate, 1.0], np.float32))
            self.trainer.update_parameters(outs[:-2])
            self.last_flat = None
            return float(outs[-2].numpy[0]) if read_loss else None
        phases = getattr(self, "phase_log", None)
        if phases is not None:
            tf.cuda_synchronize()
            t0 = time.perf_counter()
        pool, seed, flat, state = self.grad_program(self.trainer, batch_ids, np.array([self.fire_rate], np.float32))
        self.trainer.pool = pool
        self.model.seed = seed
        if phases is not None:
            tf.cuda_synchronize()
            t1 = time.perf_counter()
        # Measured on the 8-GPU box (profiles/README.md, r01c): with the NCCL allreduce enqueued behind the ~800 pending kernels of the
        # grad program a step takes 105-110 ms; with the stream drained before and after the exchange the same step takes 75.8 ms
        # (grad 75.1 + exchange 0.5 + apply 0.15), which is also what 8 independent single-GPU processes take.  The host sync is free
        # here (every step already starts with one: the batch ids are uploaded), so the exchange is bracketed by default;
        # self.diag ("async" / "before" / "after" / "skip") overrides it for diagnosis.
        # The peer-memory exchange (default) is one kernel of ours on the runtime stream: no NCCL launch path, nothing to drain.
        method = getattr(self, "method", "nccl")
        drain = self.exchange is None and method == "nccl" and os.environ.get("TFCUDA_DP_SYNC", "1") != "0"
        diag = getattr(self, "diag", "") or ("before+after" if drain else "async")
        if self.world > 1 and "skip" not in diag:
            if "before" in diag:
                tf.cuda_synchronize()
            if self.exchange is not None:
                flat = self.exchange(flat, self.world)
            elif method == "peer":
                tf.cuda_allreduce(flat, 1.0 / self.world, "peer")
            else:
                tf.cuda_allreduce(flat, 1.0 / self.world, "nccl")
            if "after" in diag:
                tf.cuda_synchronize()
        if phases is not None:
            tf.cuda_synchronize()
            t2 = time.perf_counter()
        new_params = self.apply_program(self.opt, flat, np.array([lr], np.float32))
        self.opt.update_parameters(new_params)
        if phases is not None:
            tf.cuda_synchronize()
            t3 = time.perf_counter()
            phases.append([(t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3])
        self.last_flat, self.last_state = flat, state
        if read_loss:
            return float(np.array(flat.numpy)[-1])
        return None

    def parameters_numpy(self):
        return [np.array(p.numpy) for p in self.opt.parameters()]

    def replicated_numpy(self):
        """What must be bit-identical on every rank: the model's trainable tensors (the per-rank RNG seed is NOT: it is offset by rank)."""
        return [np.array(getattr(self.model, name).numpy) for name in ("fc1", "fc1_bias", "fc2", "fc2_bias")]


# ----------------------------------------------------------------------------------------------------------------------
# communicator bootstrap: the NCCL unique id travels through torch.distributed's rendezvous store
# ----------------------------------------------------------------------------------------------------------------------
def init_comm(tf, rank, world):
    """Sets up the gradient exchange: the NCCL communicator (always: fallback and large payloads) and, unless
    TFCUDA_DP_EXCHANGE=nccl, the one-shot peer-memory exchange (csrc/comm.cu: every rank maps every peer's exchange buffer through
    CUDA IPC handles, shared here over the control-plane process group).  Returns the exchange in use: "peer" | "nccl"."""
    if world == 1:
        return "none"
    import torch.distributed as dist
    box = [tf.cuda_comm_unique_id()] if rank == 0 else [None]
    dist.broadcast_object_list(box, src=0)
    tf.cuda_comm_init(box[0], rank, world)
    if os.environ.get("TFCUDA_DP_EXCHANGE", "peer") == "nccl" or not hasattr(tf, "cuda_peer_export"):
        return "nccl"
    try:
        handle = tf.cuda_peer_export()
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        tf.cuda_peer_init(handles, rank, world)
        ok = 1
    except RuntimeError as e:  # no peer access between these devices: every rank must take the same decision
        sys.stderr.write(f"[nca_dp rank {rank}] peer-memory exchange unavailable ({e}); using NCCL\n")
        ok = 0
    votes = [None] * world
    dist.all_gather_object(votes, ok)
    if not all(votes):
        dist.barrier()
        return "nccl"
    # every rank mapped its peers.  Self-check before the exchange is trusted: the same vector through the peer kernel and through
    # ncclAllReduce (this also gives the NCCL communicator its first collective)
    probe = (np.arange(1024, dtype=np.float32) * 0.25 + rank).astype(np.float32)
    a, b = tf.cuda_tensor(probe), tf.cuda_tensor(probe)
    tf.cuda_allreduce(a, 1.0 / world, "peer")
    tf.cuda_allreduce(b, 1.0 / world, "nccl")
    want = (np.arange(1024, dtype=np.float64) * 0.25 + (world - 1) / 2.0)
    ok = int(np.allclose(tf.cuda_numpy(a), want, rtol=1e-6) and np.allclose(tf.cuda_numpy(a), tf.cuda_numpy(b), rtol=1e-6))
    dist.all_gather_object(votes, ok)
    dist.barrier()
    return "peer" if all(votes) else "nccl"


def _timed_iterations(tf, tr, steps, dist=None):
    """K iterations bracketed by device syncs (and a barrier across ranks); returns (device ms by CUDA events, host ms spent enqueuing)."""
    tf.cuda_synchronize()
    if dist is not None:
        dist.barrier()
    tf.cuda_timer_begin()
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.step()
    host_ms = (time.perf_counter() - t0) * 1e3
    ms = tf.cuda_timer_end()
    tf.cuda_synchronize()
    return ms, host_ms


def bench_main(args):
    """bench.py --workload nca (the default under torchrun with N > 1): K training iterations of the data-parallel NCA config, strong
    scaling (global batch fixed, split across ranks); value = samples/s over all ranks, time = CUDA events on every rank's stream,
    max over ranks.  The line also carries what the scaling is measured against, taken in the same job on rank 0's GPU while the other
    ranks wait: `weak.alone_ms_per_step` (the same per-rank program without peers) and `single_gpu` (the whole global batch on one GPU)."""
    import faulthandler
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # a hung collective must not hold the GPUs until the lease ends: dump every thread's stack and exit
    faulthandler.dump_traceback_later(int(os.environ.get("TFCUDA_BENCH_DEADLINE", "800")), exit=True)
    dist = None
    if world > 1:
        import datetime
        import torch  # before TensorFrost
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("gloo", timeout=datetime.timedelta(seconds=900))  # control plane only: barriers, handle exchange, max-over-ranks
    import tensorfrost_b200
    weak = bool(getattr(args, "nca_weak", False))
    if weak:
        # weak scaling: the per-GPU batch and pool shard stay at --nca-batch / --nca-pool, the global ones grow with N
        args.nca_batch *= world
        args.nca_pool *= world
    sys.path.insert(0, REPO_ROOT)
    from bench import ClockSampler, quiet_stdout  # nvidia-smi clocks / throttle reasons of THIS rank's GPU during the timed region
    matmul = getattr(args, "nca_matmul", "tf32")
    with quiet_stdout():
        tf = tensorfrost_b200.load((os.environ.get("TFCUDA_KERNEL_OPTIONS", "") + " --tf-matmul=" + matmul).strip())
        method = init_comm(tf, rank, world)
        t0 = time.perf_counter()
        tr = NcaTrainer(tf, global_batch=args.nca_batch, grid=args.nca_grid, pool_size=args.nca_pool, train_steps=args.nca_steps,
                        rank=rank, world=world, mono=args.nca_mono)
        tr.method = method
        build_s = time.perf_counter() - t0
        emitter = None
        try:  # how the per-rank programs were emitted: kernels, and how many carry several elements ("lanes") per thread
            texts = [k[0][1] + k[0][2] for k in tf.get_all_generated_kernels()]
            texts = [t for t in texts if "__global__" in t]
            emitter = {"kernels": len(texts), "with_lanes": sum("lanes per thread" in t for t in texts),
                       "lanes_per_thread": int(os.environ.get("TFCUDA_COARSEN", "4") or 0)}
        except Exception:  # noqa: BLE001
            pass
        # the first ~25 iterations of a fresh process run up to 10 % slower than the steady state (measured at 8 GPUs: 70.0 ms averaged over
        # iterations 6-25, 63.2 ms afterwards; the reference pool above the runtime is still adapting its expiry times): warm up at least 12
        # iterations, the same for the two single-GPU references below, and report the number actually used
        warm = max(args.warmup, 12)
        for _ in range(warm):
            tr.step()
        sampler = ClockSampler(local)
        sampler.start()
        launches0 = tf.cuda_launch_count()
        driver_calls0 = tf.cuda_pool_driver_calls()
        ms, host_issue_ms = _timed_iterations(tf, tr, args.steps, dist)
        clocks = sampler.stop()
        launches = tf.cuda_launch_count() - launches0
        driver_calls = tf.cuda_pool_driver_calls() - driver_calls0
        graph = tf.cuda_graph_stats() if hasattr(tf, "cuda_graph_stats") else None
        loss = tr.step(read_loss=True)
        # end to end: the same iterations through the public API with HOST data every step - the batch indices and the step's scalars go
        # up (pageable numpy arrays), the exchanged [gradients, loss] vector comes back and the loss is read on the host (which also
        # synchronises every iteration); wall clock between device syncs, max over ranks
        e2e_steps = max(3, args.steps // 2)
        tf.cuda_synchronize()
        if dist is not None:
            dist.barrier()
        t_e2e = time.perf_counter()
        for _ in range(e2e_steps):
            loss = tr.step(read_loss=True)
        tf.cuda_synchronize()
        e2e_s = time.perf_counter() - t_e2e
        # replicated state must be bit-identical on every rank (the optimizer step is replicated, nothing is broadcast)
        digest = None
        if not args.nca_mono:
            import hashlib
            h = hashlib.sha256()
            for p in tr.replicated_numpy():
                h.update(np.ascontiguousarray(p).tobytes())
            digest = h.hexdigest()
        phase_ms = None
        if os.environ.get("TFCUDA_NCA_PHASES"):
            # diagnosis: three more steps with a device sync after each phase (grad program / gradient exchange / apply program)
            tr.phase_log = []
            for _ in range(3):
                tr.step()
            phase_ms = tr.phase_log
            tr.phase_log = None
        diag_ms = None
        if os.environ.get("TFCUDA_NCA_DIAG"):
            # diagnosis of the exchange: the same 3 steps with a device sync before / after the allreduce, or without the allreduce
            diag_ms = {}
            for mode in ("async", "before", "after", "before+after", "skip", "async"):
                tr.diag = mode
                tr.step()
                d_ms, _ = _timed_iterations(tf, tr, 3, dist)
                diag_ms[mode + ("#2" if mode in diag_ms else "")] = d_ms / 3
            tr.diag = ""
        top = None
        if getattr(args, "nca_profile", False):
            # EVERY rank takes the profiled step (the exchange is a collective: a rank-0-only step would wait for its peers forever)
            tf.cuda_profile_reset()
            tf.cuda_profile_enable(True)
            tr.step()
            tf.cuda_profile_enable(False)
            if rank == 0:
                recs = sorted(tf.cuda_profile_records(), key=lambda r: -r["total_ms"])
                total = sum(r["total_ms"] for r in recs)
                top = [{"name": r["name"], "launches": r["launches"], "ms": round(r["total_ms"], 3), "share": round(r["total_ms"] / total, 4),
                        "gbs": round(r["bytes"] / max(r["total_ms"], 1e-9) / 1e6, 1)} for r in recs[:24]]
                dump = os.environ.get("TFCUDA_PROFILE_DUMP")
                if dump:
                    with open(dump, "w") as f:
                        json.dump(recs, f)
                top.append({"name": "TOTAL", "launches": sum(r["launches"] for r in recs), "ms": round(total, 3), "gb": round(sum(r["bytes"] for r in recs) / 1e9, 2)})
        per_rank_ms, digests = [ms / args.steps], [digest]
        alone = single = None
        if dist is not None:
            every = [None] * world
            dist.all_gather_object(every, (ms, digest, e2e_s))
            per_rank_ms = [e[0] / args.steps for e in every]
            digests = [e[1] for e in every]
            ms = max(e[0] for e in every)
            e2e_s = max(e[2] for e in every)
            if not getattr(args, "no_single", False):
                # references for the scaling numbers, on rank 0's GPU while the peers idle at the barrier below:
                if rank == 0:
                    # (weak) the SAME per-rank program and batch without peers: T(1 GPU, B/N) against T(N GPUs, B/N each)
                    tr.diag = "skip"
                    for _ in range(2):
                        tr.step()
                    a_ms, _ = _timed_iterations(tf, tr, args.steps)
                    tr.diag = ""
                    alone = {"alone_ms_per_step": a_ms / args.steps, "per_gpu_batch": args.nca_batch // world,
                             "note": "rank 0's per-rank program without the exchange while the other GPUs idle"}
                    # (strong) the whole global batch on ONE GPU: T(1 GPU, B)
                    t1 = time.perf_counter()
                    one = NcaTrainer(tf, global_batch=args.nca_batch, grid=args.nca_grid, pool_size=args.nca_pool, train_steps=args.nca_steps,
                                     rank=0, world=1)
                    for _ in range(warm):
                        one.step()
                    s_ms, _ = _timed_iterations(tf, one, max(3, args.steps // 2))
                    s_ms /= max(3, args.steps // 2)
                    single = {"ms_per_step": s_ms, "samples_per_s": args.nca_batch / (s_ms / 1e3), "global_batch": args.nca_batch,
                              "build_seconds": time.perf_counter() - t1, "note": "the whole global batch on rank 0's GPU alone, same job, same box"}
                    del one
                dist.barrier()
    if rank == 0:
        samples = args.nca_batch * args.steps
        line = {
            "metric": "NCA training samples/s", "value": samples / (ms / 1e3), "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"NCA training (examples/ML/NCA, BASELINE configs[4]), global batch {args.nca_batch} of {args.nca_grid}x{args.nca_grid}x12, "
                                   f"{args.nca_steps} CA steps, pool {args.nca_pool}", "parallelism": f"dp{world}",
                       "per_rank_batch": args.nca_batch // world,
                       "exchange": {"peer": "one-shot allreduce kernel over NVLink peer memory (7821 fp32, rank-order sum) on the runtime stream",
                                    "nccl": "ncclAllReduce(sum) of 7821 fp32 + scale, stream drained around it", "none": "none (1 GPU)"}[method],
                       "program": "reference single program" if args.nca_mono else "grad program -> allreduce -> apply program",
                       "matmul": {"tf32": "tf.initialize(tf.cuda, '--tf-matmul=tf32'): one tcgen05 kind::tf32 product per matmul (1e-3 class)",
                                  "3xtf32": "backend default: 3xTF32 split products (fp32-accurate)", "fp32": "FFMA kernel"}[matmul]},
            "gpu_launches": int(launches), "loss_after": loss, "build_seconds": build_s, "emitter": emitter, "clocks": clocks,
            "host_issue_ms_per_step": host_issue_ms / args.steps, "device_alloc_calls_per_step": driver_calls / args.steps, "graph": graph,
            "host_cores": os.cpu_count(),
            "e2e": {"value": args.nca_batch * e2e_steps / e2e_s, "unit": "samples/s", "ms_per_step": e2e_s / e2e_steps * 1e3, "steps": e2e_steps,
                    "h2d_bytes_per_step": world * (4 * (args.nca_batch // world) + 8), "d2h_bytes_per_step": world * 4 * 7821,
                    "path": "NcaTrainer.step(read_loss=True): numpy batch ids + scalars in, exchanged gradient vector + loss out, every iteration"},
            "verify": {"ok": bool(np.isfinite(loss)) and len(set(digests)) == 1, "loss_finite": bool(np.isfinite(loss)),
                       "parameters_bit_identical_across_ranks": len(set(digests)) == 1, "parity": "tests/test_nca_gpu.py (split and mono step vs reference golden)"},
        }
        if phase_ms is not None:
            line["phase_ms_grad_exchange_apply"] = phase_ms
        if diag_ms is not None:
            line["diag_ms_per_step"] = diag_ms
        if top is not None:
            line["top_kernels"] = top
        if world > 1:
            line["per_rank_ms_per_step"] = per_rank_ms
            line["weak"] = alone
            line["single_gpu"] = single
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    faulthandler.cancel_dump_traceback_later()
    return 0
