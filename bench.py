#!/usr/bin/env python3
"""bench.py — benchmark of the B200 CUDA backend for TensorFrost programs.

N = 1 (the headline, BASELINE.json configs[1]): one step of the 2-D Eulerian fluid simulation on a 2048 x 2048 fp32 grid
(the reference's examples/Simulation/fluid_simulation.ipynb program, 15 fused kernels / 43 dispatches per step), traced by the
unchanged TensorFrost frontend and executed by the CUDA backend (emitter + libtfcuda.so).
  metric = fused-kernel HBM GB/s = ALGORITHMIC bytes of a step / device time of a step; the algorithmic bytes are the sum over
  the step's dispatches of the size of every tensor bound to the dispatch, each counted once.
  The same line carries BASELINE.json's other configs (`extra`: radix sort 2^28, n-body 262144, row reductions / scan / matmul
  8192^2 — each with roofline, output verification, end-to-end number and CPU baseline) and the 1-GPU point of the
  data-parallel NCA config (`nca_dp`).

N > 1 (launched by torchrun): the only config that shards — data-parallel NCA training, global batch 256 of
128x128x12, 25 CA steps, gradient allreduce; strong scaling, metric = samples/s.  The line also holds what the scaling is
measured against, taken in the same job: rank 0 alone on the whole batch (`single_gpu`) and on its own shard (`weak`).
`--workload fluid` under torchrun runs N independent replicas of the fluid step instead ("replicas only").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size 2048] [--no-extra] [--workload fluid|nca]

--impl reference: the reference's own C++/OpenMP backend (oracle/_ref) on the host cores, same program and size, all host threads.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fused-kernel HBM GB/s"


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class quiet_stdout:
    """The reference prints import / compile chatter on stdout (TensorProgram properties, temp file names): keep stdout for the
    ONE JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.devnull = os.open(os.devnull, os.O_WRONLY)
        self.saved = os.dup(1)
        os.dup2(self.devnull, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        os.close(self.devnull)


# ------------------------------------------------------------------------------------------------------------------
# the workload description both arms print (identical dicts: the driver compares them)
# ------------------------------------------------------------------------------------------------------------------
def fluid_step_bytes(n):
    """Algorithmic bytes of one fluid step = sum over its 43 dispatches of every bound tensor, once.  With
    P = 4 n^2 bytes per full-resolution field: k0 6P, k1 5P, k2 3P, k3 1.5P, 8 sweeps at n/2 of 0.75P, k6 1.5P, k7 0.375P, 24 sweeps
    at n/4 of 0.1875P, k10 0.5625P, k11 2.25P, k12 3P, k13 3P, k14 12P = 48.6875 P, plus mouse (20 B) and params (24 B, bound twice).
    The CUDA arm counts the same quantity live in the runtime's profiler and reports it next to this closed form."""
    return int(48.6875 * 4 * n * n) + 68


def fluid_config(n):
    return {"workload": f"fluid_simulation {n}x{n} fp32 (BASELINE configs[1]), 1 step = 43 dispatches of 15 fused kernels",
            "bytes_per_step": fluid_step_bytes(n),
            "l2": "per-step working set (~20 fields x 16.8 MB) exceeds the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own C++/OpenMP backend on the host cores
# ------------------------------------------------------------------------------------------------------------------
def load_reference():
    """import the UNMODIFIED reference module (oracle/_ref) with every host thread available to its OpenMP kernels (torchrun
    sets OMP_NUM_THREADS=1 for its children)."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref, "TensorFrost")):
        return None
    os.environ["OMP_NUM_THREADS"] = str(host_cores())
    sys.path.insert(0, ref)
    import TensorFrost as tf
    return tf


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    n = args.size
    with quiet_stdout():
        tf = load_reference()
        if tf is not None:
            tf.initialize(tf.cpu)  # the reference's default flags: -O3 -ffast-math -fopenmp
            from tensorfrost_b200 import workloads
            if args.workload == "extras":
                out = reference_extras(tf, workloads, args)
            else:
                fluid = workloads.load_fluid(tf, n, n)
                if args.dump_state:
                    # parity leg of the CUDA arm: `verify_steps` steps from rest, fields dumped for comparison
                    state = [tf.tensor(a) for a in workloads.fluid_inputs(n, n)]
                    for _ in range(args.verify_steps):
                        state, _ = workloads.fluid_step(fluid, state)
                    np.savez(args.dump_state, **{k: np.array(t.numpy) for k, t in zip(("vx", "vy", "pressure", "density"), state[:4])})
                state = [tf.tensor(a) for a in workloads.fluid_inputs(n, n)]
                for _ in range(args.warmup):
                    state, _ = workloads.fluid_step(fluid, state)
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    state, _ = workloads.fluid_step(fluid, state)
                dt = time.perf_counter() - t0
    if tf is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (reference module built by oracle/build_ref.sh) is missing"}))
        return 0
    if args.workload == "extras":
        print(json.dumps({"impl": "reference", "extras": out, "cores": host_cores()}))
        return 0
    step_bytes = fluid_step_bytes(n)
    value = step_bytes * args.steps / dt / 1e9
    cores = host_cores()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": fluid_config(n),
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores, "kind": "reference",
                         "sample": f"{args.steps} steps of the {n}x{n} fluid program on tf.cpu (-O3 -ffast-math -fopenmp, {cores} OpenMP threads)"},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def reference_extras(tf, workloads, args):
    """CPU baselines of the other configs on the reference's C++/OpenMP backend, each on a bounded sample (stated) so the whole leg
    takes well under a minute."""
    import numpy as np
    rng = np.random.default_rng(0)
    out = {}

    def timed(fn, reps):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps

    def guard(name, fn):
        try:
            fn()
        except Exception as e:  # noqa: BLE001
            out[name] = {"value": None, "error": f"{type(e).__name__}: {e}"[:200]}

    def sort():
        n = 1 << 20  # BASELINE configs[0]: tests/sorting_test.py radix sort of 2^20 uint32 on the C++/OpenMP backend
        keys = tf.tensor(rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32))
        vals = tf.tensor(np.arange(n, dtype=np.uint32))
        prog = workloads.compile_sort(tf, with_values=True)
        s = timed(lambda: prog(keys, vals), 3)
        out["radix_sort"] = {"value": n / s / 1e9, "unit": "Gkeys/s", "sample": "tf.sort.radix(keys, values), 2^20 uint32 pairs (configs[0])"}

    def nbody():
        nb = 8192
        x = tf.tensor((5.0 * rng.standard_normal((nb, 3))).astype(np.float32))
        v = tf.tensor(np.zeros((nb, 3), np.float32))
        prog = workloads.compile_nbody(tf)
        s = timed(lambda: prog(x, v), 2)
        out["nbody"] = {"value": nb * nb / s / 1e9, "unit": "Ginteractions/s", "sample": f"n_body program at {nb} bodies (work scales with N^2: 1/1024 of the 262144-body step)"}

    def reductions():
        m = 4096
        a = tf.tensor(rng.random((m, m), dtype=np.float32))
        prog = workloads.compile_row_reductions(tf, m)
        s = timed(lambda: prog(a), 3)
        out["row_reductions"] = {"value": 4 * m * m * 4 / s / 1e9, "unit": "GB/s", "sample": f"tf.sum/max/mean/norm over the rows of a {m}x{m} matrix (4 reads of A counted)"}

    def matmul():
        m = 1024
        a = tf.tensor(rng.random((m, m), dtype=np.float32))
        b = tf.tensor(rng.random((m, m), dtype=np.float32))
        prog = workloads.compile_matmul(tf)
        s = timed(lambda: prog(a, b), 2)
        out["matmul"] = {"value": 2.0 * m ** 3 / s / 1e12, "unit": "TFLOP/s", "sample": f"a @ b at {m}^3 (1/512 of the 8192^3 product)"}

    guard("radix_sort", sort)
    guard("nbody", nbody)
    guard("row_reductions", reductions)
    guard("matmul", matmul)
    return out


# ------------------------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------------------------
def dist_setup(n_gpus):
    """torch.distributed only for the barrier and the max-over-ranks of the measured time (N > 1)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None, 0, 1
    import torch  # must be imported before TensorFrost in this image
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return dist, dist.get_rank(), world


def max_over_ranks(dist, value):
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(dist, tf):
    tf.cuda_synchronize()
    if dist is not None:
        import torch
        dist.barrier()
        torch.cuda.synchronize()


NCU_SUMMARIES = ("r02_fluid_ncu_full_summary.csv", "r01b_fluid_ncu_full_summary.csv")


def _ncu_rows():
    import csv
    for name in NCU_SUMMARIES:
        path = os.path.join(ROOT, "profiles", name)
        try:
            rows = list(csv.reader(open(path)))
            return name, rows, {h: i for i, h in enumerate(rows[0])}
        except (OSError, IndexError):
            continue
    return None, None, None


def ncu_traffic(kernel_name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of an emitted fluid kernel, from the committed `ncu --set full` capture
    of this same command (profiles/README.md); None when the capture does not hold that kernel."""
    name, rows, col = _ncu_rows()
    if rows is None:
        return None, None
    try:
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        vals = []
        for r in rows[2:]:
            if r[col["Kernel Name"]] == kernel_name:
                rd = float(r[col["dram__bytes_read.sum"]]) * unit[rows[1][col["dram__bytes_read.sum"]]]
                wr = float(r[col["dram__bytes_write.sum"]]) * unit[rows[1][col["dram__bytes_write.sum"]]]
                vals.append(rd + wr)
        return (sum(vals) / len(vals) if vals else None), name
    except (KeyError, ValueError, IndexError):
        return None, name


def ncu_limiter(kernel_name):
    """What the same capture says limits that kernel: issue-slot, DRAM and L1 utilisation (percent of peak)."""
    _name, rows, col = _ncu_rows()
    if rows is None:
        return None
    try:
        for r in rows[2:]:
            if r[col["Kernel Name"]] == kernel_name:
                return {"issue_active_pct": float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
                        "dram_pct": float(r[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
                        "l1tex_pct": float(r[col["l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]]),
                        "warp_instructions": float(r[col["smsp__inst_executed.sum"]])}
    except (KeyError, ValueError, IndexError):
        pass
    return None


def dominant(records, steps):
    """Pick the kernel with the largest share of device time; return its roofline fields."""
    total = sum(r["total_ms"] for r in records) or 1.0
    top = max(records, key=lambda r: r["total_ms"])
    per_launch_ms = top["total_ms"] / max(top["launches"], 1)
    per_launch_bytes = top["bytes"] / max(top["launches"], 1)
    return top, per_launch_ms, per_launch_bytes, top["total_ms"] / total


def rel_err(got, want, floor):
    """max over elements of |got - want| / max(|want|, floor): elementwise relative error with an absolute floor."""
    import numpy as np
    g, w = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float(np.max(np.abs(g - w) / np.maximum(np.abs(w), floor))) if g.size else 0.0


def verify_fluid(tf, fluid, workloads, args):
    """Parity of the TIMED program object at the TIMED size: the parity scenario of tests/test_zz_fluid_gpu.py (`verify_steps` steps
    from rest with a moving source, outputs fed back) on the CUDA backend against the reference's C++/OpenMP backend (oracle/_ref,
    strict flags: no fast-math) run in a subprocess on this box's host with the same inputs.  Two measures per field: norm-wise
    max|got-ref| / max|ref|, and element-wise |got-ref| / max(|ref|, 0.01 max|ref|)."""
    import tempfile
    import numpy as np
    n = args.size
    if not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "TensorFrost")):
        return {"ok": None, "skipped": "oracle/_ref missing"}
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "ref_state.npz")
        cmd = [sys.executable, os.path.join(ROOT, "tests", "golden", "make_golden_fluid.py"), "run", "strict", path, str(n), str(n), str(args.verify_steps)]
        env = dict(os.environ, OMP_NUM_THREADS=str(host_cores()))
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=d, env=env)
        if r.returncode != 0 or not os.path.exists(path):
            return {"ok": None, "skipped": "reference run failed: " + r.stderr[-200:]}
        want = dict(np.load(path))
    got = workloads.fluid_parity_run(tf, n, n, args.verify_steps, program=fluid)
    names = ["vx", "vy", "pressure", "density", "div", "canvas"]
    norm, elem = {}, {}
    for k, g in zip(names, got):
        w = want[k].astype(np.float64)
        scale = max(float(np.max(np.abs(w))), 1e-30)
        norm[k] = float(np.max(np.abs(g.astype(np.float64) - w))) / scale
        elem[k] = rel_err(g, w, 0.01 * scale)
    bars = {"vx": 2e-5, "vy": 2e-5, "pressure": 2e-5, "density": 2e-5, "div": 1e-4, "canvas": 2e-5}  # tests/test_zz_fluid_gpu.py
    ok = all(norm[k] <= bars[k] for k in names) and all(np.isfinite(list(elem.values())))
    return {"ok": bool(ok), "steps": args.verify_steps, "max_err_over_max_ref": norm, "max_elementwise_rel_err_floor_1pct": elem, "bars_normwise": bars,
            "against": "reference C++/OpenMP backend (oracle/_ref, -O3 -fopenmp, no fast-math), same inputs, this box's host"}


def emitter_stats(tf):
    """How the timed program was emitted: kernels, and how many of them carry several elements ("lanes") per thread."""
    try:
        texts = [k[0][1] + k[0][2] for k in tf.get_all_generated_kernels()]
    except Exception:  # noqa: BLE001
        return None
    texts = [t for t in texts if "__global__" in t]
    return {"kernels": len(texts), "with_lanes": sum("lanes per thread" in t for t in texts),
            "lanes_per_thread": int(os.environ.get("TFCUDA_COARSEN", "4") or 0), "range_fact": os.environ.get("TFCUDA_ASSUME", "1") != "0"}


FLUID_WARMUP = 8


def fluid_warmup(args):
    return max(args.warmup, FLUID_WARMUP)


def bench_fluid(tf, dist, rank, world, args, peaks):
    import numpy as np
    from tensorfrost_b200 import workloads
    n = args.size
    fluid = workloads.load_fluid(tf, n, n)
    emitter = emitter_stats(tf) if hasattr(tf, "get_all_generated_kernels") else None
    host_inputs = workloads.fluid_inputs(n, n)
    state = [tf.cuda_tensor(a) for a in host_inputs]
    # warm-up: at least FLUID_WARMUP steps (reported in the line).  The launch recorder replays the step as a CUDA graph only once the
    # chain has PROVEN to repeat: the pool alternates between two address sets, each is launched eagerly until seen twice and then
    # instantiated, so the first 5 executions are eager and carry two graph instantiations (~0.5 ms of host time each).  With 3 warm-up
    # steps those one-time costs sat inside the timed region (round 2: 0.366 ms/step reported, ~0.30 in the steady state).
    for _ in range(fluid_warmup(args)):
        state, _ = workloads.fluid_step(fluid, state)
    # ---- timed region: K steps, inputs resident in HBM, CUDA events on the launching stream --------------------
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    barrier(dist, tf)
    launches0 = tf.cuda_launch_count()
    tf.cuda_timer_begin()
    for _ in range(args.steps):
        state, _ = workloads.fluid_step(fluid, state)
    ms = tf.cuda_timer_end()
    barrier(dist, tf)
    launches = tf.cuda_launch_count() - launches0
    clocks = sampler.stop()
    ms = max_over_ranks(dist, ms)
    graph = tf.cuda_graph_stats() if hasattr(tf, "cuda_graph_stats") else None
    # ---- per-kernel profile of the same K steps (event pair per launch, eager launches) -> bytes per step and the dominant kernel ----
    tf.cuda_profile_reset()
    tf.cuda_profile_enable(True)
    for _ in range(args.steps):
        state, _ = workloads.fluid_step(fluid, state)
    tf.cuda_profile_enable(False)
    records = tf.cuda_profile_records()
    counted_bytes = sum(r["bytes"] for r in records) / args.steps
    step_bytes = fluid_step_bytes(n)
    top, top_ms, top_bytes, share = dominant(records, args.steps)
    value = world * step_bytes * args.steps / (ms / 1e3) / 1e9
    achieved = top_bytes / (top_ms / 1e3) / 1e9 if top_ms > 0 else 0.0
    traffic, traffic_file = ncu_traffic(top["name"]) if n == 2048 else (None, None)
    kernel_sum_ms = sum(r["total_ms"] for r in records) / args.steps
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": traffic, "traffic_source": f"profiles/{traffic_file} (ncu --set full, same command)" if traffic_file else None,
                "kernel": top["name"], "ncu": ncu_limiter(top["name"]) if n == 2048 else None, "share_of_step": share, "launch_ms": top_ms,
                "algorithmic_bytes_per_launch": top_bytes, "peak_source": peaks["source"],
                "whole_step": {"achieved": value / world, "frac": value / world / peaks["hbm_gbs"], "sum_of_kernel_ms": kernel_sum_ms,
                               "note": "algorithmic bytes of the step / device time of the step (timed loop); sum_of_kernel_ms = the 43 kernels timed one by one"}}
    # ---- e2e: the same step through the public API with HOST buffers: every step uploads its 4 input fields from page-locked host
    # memory and downloads its 4 result fields into page-locked host memory.  Uploads, kernels and downloads run on three streams
    # (tf.cuda_upload_async / cuda_wait_uploads / cuda_download_async): step k+1's upload overlaps step k's kernels and download. ----
    step_shape = [n, n]
    pinned_in = [tf.cuda_pinned_array(step_shape, "float32") for _ in range(4)]
    pinned_out = [[tf.cuda_pinned_array(step_shape, "float32") for _ in range(4)] for _ in range(2)]
    for p, a in zip(pinned_in, host_inputs[:4]):
        p[...] = a
    sets = [[tf.cuda_tensor(a) for a in host_inputs] for _ in range(2)]
    pipelined = hasattr(tf, "cuda_upload_async")

    def e2e_loop(steps):
        if pipelined:
            for k in range(4):
                tf.cuda_upload_async(sets[0][k], pinned_in[k])
            for s in range(steps):
                cur, nxt = sets[s % 2], sets[(s + 1) % 2]
                tf.cuda_wait_uploads()                      # kernels below see upload(s)
                if s + 1 < steps:
                    for k in range(4):
                        tf.cuda_upload_async(nxt[k], pinned_in[k])   # upload(s+1): waits only for step s-1 (last user of that set)
                out_state, _ = workloads.fluid_step(fluid, cur)
                for k in range(4):
                    tf.cuda_download_async(out_state[k], pinned_out[s % 2][k])
            tf.cuda_copy_sync()
            tf.cuda_synchronize()
        else:
            e2e_state = sets[0]
            for _ in range(steps):
                for k in range(4):
                    tf.cuda_upload(e2e_state[k], pinned_in[k])
                e2e_state, _ = workloads.fluid_step(fluid, e2e_state)
                for k in range(4):
                    tf.cuda_download(e2e_state[k], pinned_out[0][k])
            tf.cuda_synchronize()

    e2e_loop(fluid_warmup(args))  # warm-up: the pipeline keeps more tensors alive than the resident loop did (new device blocks, new graphs)
    barrier(dist, tf)
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    e2e_s = max_over_ranks(dist, time.perf_counter() - t0)
    e2e = {"value": world * step_bytes * args.steps / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": 4 * n * n * 4,
           "d2h_bytes_per_step": 4 * n * n * 4, "ms_per_step": e2e_s / args.steps * 1e3,
           "path": "page-locked host arrays <-> device on copy streams overlapping the kernels (tf.cuda_upload_async / cuda_download_async)"
                   if pipelined else "page-locked host arrays <-> device, blocking copies (tf.cuda_upload / cuda_download)"}
    verify = None
    if rank == 0 and not getattr(args, "no_verify", False):
        try:
            verify = verify_fluid(tf, fluid, workloads, args)
        except Exception as e:  # noqa: BLE001
            verify = {"ok": None, "skipped": f"{type(e).__name__}: {e}"[:200]}
    return {"value": value, "ms": ms, "launches": launches, "clocks": clocks, "roofline": roofline, "e2e": e2e, "step_bytes": step_bytes,
            "counted_bytes": counted_bytes, "graph": graph, "verify": verify, "emitter": emitter,
            "records": sorted(records, key=lambda r: -r["total_ms"])[:16]}


def time_call(tf, fn, iters, warm=3):
    for _ in range(warm):
        fn()
    tf.cuda_synchronize()
    tf.cuda_timer_begin()
    for _ in range(iters):
        fn()
    return tf.cuda_timer_end() / iters


def bench_extra(tf, peaks, quick, only=None):
    """BASELINE.json's other metrics at the configs' sizes, each against its own roofline, each with its timed output VERIFIED
    (sort: sortedness + stability + permutation; n-body / matmul / reductions: float64 on sampled rows) and an end-to-end figure
    (host buffers in, host buffers out).  Every section is independent: a failure is recorded under its name."""
    import numpy as np
    from tensorfrost_b200 import workloads
    out = {}
    rng = np.random.default_rng(0)
    hbm = peaks["hbm_gbs"]
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    tf32_peak = peaks["bf16_tflops"] / 2

    def section(name, fn):
        if only is not None and name not in only:
            return
        try:
            fn()
        except Exception as e:  # noqa: BLE001
            out[name + "_error"] = f"{type(e).__name__}: {e}"[:300]

    def sort_section():
        # radix sort, 2^28 uint32 keys (keys-only: 36 B/key algorithmic; key+value: 68 B/pair)
        n = 1 << (24 if quick else 28)
        host_keys = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
        keys = tf.cuda_tensor(host_keys)
        ms = time_call(tf, lambda: tf.cuda_radix_sort(keys), 5)
        sorted_keys = tf.cuda_numpy(tf.cuda_radix_sort(keys))
        ok_keys = bool(np.all(sorted_keys[1:] >= sorted_keys[:-1])) and int(np.bitwise_xor.reduce(sorted_keys)) == int(np.bitwise_xor.reduce(host_keys)) \
            and int(sorted_keys.sum(dtype=np.uint64)) == int(host_keys.sum(dtype=np.uint64))
        out["radix_sort_keys"] = {"n": n, "ms": ms, "gkeys_per_s": n / ms / 1e6, "roofline": {"bound": "hbm", "achieved": 36.0 * n / ms / 1e6, "peak": hbm,
                                  "unit": "GB/s", "frac": 36.0 * n / ms / 1e6 / hbm},
                                  "verify": {"ok": ok_keys, "how": "ascending; xor and sum of keys preserved (checksum of checksums)"}}
        del sorted_keys
        vals = tf.cuda_tensor(np.arange(n, dtype=np.uint32))
        ms = time_call(tf, lambda: tf.cuda_radix_sort(keys, vals), 5)
        k2, v2 = tf.cuda_radix_sort(keys, vals)
        k2, v2 = tf.cuda_numpy(k2), tf.cuda_numpy(v2)
        asc = k2[1:] >= k2[:-1]
        stable = bool(np.all(asc & ((k2[1:] != k2[:-1]) | (v2[1:] > v2[:-1]))))       # equal keys keep their input order
        perm = bool(np.array_equal(host_keys[v2], k2))                                  # values are the permutation that sorts the keys
        out["radix_sort_pairs"] = {"n": n, "ms": ms, "gkeys_per_s": n / ms / 1e6, "roofline": {"bound": "hbm", "achieved": 68.0 * n / ms / 1e6, "peak": hbm,
                                   "unit": "GB/s", "frac": 68.0 * n / ms / 1e6 / hbm},
                                   "verify": {"ok": stable and perm, "how": "ascending, stable (ties keep input order), keys_in[values_out] == keys_out: "
                                                                               "together these pin the result to np.argsort(kind='stable')"}}
        del k2, v2, asc
        sort_prog = workloads.compile_sort(tf, with_values=True)  # tf.sort.radix inside a compiled program
        ms = time_call(tf, lambda: sort_prog(keys, vals), 5)
        out["radix_sort_pairs_program"] = {"n": n, "ms": ms, "gkeys_per_s": n / ms / 1e6, "note": "tf.sort.radix(keys, values) traced by tf.compile: one library call + output copies"}
        # end to end: keys in page-locked host memory -> device -> sort -> page-locked host memory
        pin_in = tf.cuda_pinned_array([n], "uint32")
        pin_out = tf.cuda_pinned_array([n], "uint32")
        pin_in[...] = host_keys
        dev = tf.cuda_tensor(host_keys)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            tf.cuda_upload(dev, pin_in)
            res = tf.cuda_radix_sort(dev)
            tf.cuda_download(res, pin_out)
        s = (time.perf_counter() - t0) / reps
        out["radix_sort_keys"]["e2e"] = {"value": n / s / 1e9, "unit": "Gkeys/s", "h2d_bytes_per_step": 4 * n, "d2h_bytes_per_step": 4 * n, "ms": s * 1e3}

    def nbody_section():
        # n-body, 262144 bodies: the reference PROGRAM on the drop-in path (emitter) and the hand-written library kernel
        nb = 32768 if quick else 262144
        hx = (5.0 * rng.standard_normal((nb, 3))).astype(np.float32)
        x = tf.cuda_tensor(hx)
        v = tf.cuda_tensor(np.zeros((nb, 3), np.float32))

        def check(xn, vn):
            # float64 all-pairs force on 256 sampled bodies (n-body-benchmark.py:16-34 semantics); the error is measured against the
            # summed magnitude of each body's 262144 terms (what an fp32 accumulation - the reference's own is serial fp32 - is accurate to)
            idx = rng.choice(nb, 256, replace=False)
            X = hx.astype(np.float64)
            d = X[idx, None, :] - X[None, :, :]
            d2 = (d ** 2).sum(-1) + 1e-4
            terms = -d / (d2 * np.sqrt(d2))[..., None]
            v_ref, a_ref = terms.sum(1) * 0.001, np.abs(terms).sum(1) * 0.001
            return float(np.max(np.abs(np.asarray(vn)[idx].astype(np.float64) - v_ref) / a_ref))

        ms = time_call(tf, lambda: tf.cuda_nbody_step(x, v), 3, warm=1)
        xn, vn = tf.cuda_nbody_step(x, v)
        err = check(tf.cuda_numpy(xn), tf.cuda_numpy(vn))
        out["nbody_library"] = {"bodies": nb, "ms": ms, "ginteractions_per_s": nb * nb / ms / 1e6,
                                "roofline": {"bound": "fp32", "achieved": 20.0 * nb * nb / ms / 1e9, "peak": fp32_peak, "unit": "TFLOP/s",
                                             "frac": 20.0 * nb * nb / ms / 1e9 / fp32_peak, "note": "20 flop/interaction (SURVEY §8d C3); CUDA-core bound, not HBM"},
                                "verify": {"ok": bool(err <= 1e-4), "max_err_over_summed_term_magnitudes_vs_float64_on_256_bodies": err}}
        for key, compile_fn, what in (("nbody_program", workloads.compile_nbody, "n_body (n-body-benchmark.py:16-34: broadcast differences + tf.sum)"),
                                      ("nbody_loop_program", workloads.compile_nbody_loop, "n_body_loop (n-body-benchmark.py:36-65: one thread per body, tf.loop over partners)")):
            nbody = compile_fn(tf)
            ms = time_call(tf, lambda: nbody(x, v), 2, warm=1)
            xn, vn = nbody(x, v)
            # n_body_loop's force law differs from n_body's (the example divides by (d2 + eps) * dist and multiplies by dx twice): check it
            # against its own float64 restatement
            if key == "nbody_program":
                err = check(tf.cuda_numpy(xn), tf.cuda_numpy(vn))
            else:
                idx = rng.choice(nb, 256, replace=False)
                X = hx.astype(np.float64)
                d = X[None, :, :] - X[idx, None, :]
                d2 = (d ** 2).sum(-1)
                g = -d[..., 0] / (d2 + 1e-4) / np.sqrt(d2 + 1e-4)
                terms = g[..., None] * d
                v_ref, a_ref = terms.sum(1) * 0.001, np.abs(terms).sum(1) * 0.001
                err = float(np.max(np.abs(tf.cuda_numpy(vn)[idx].astype(np.float64) - v_ref) / a_ref))
            t0 = time.perf_counter()
            xn, vn = nbody(tf.cuda_tensor(hx), v)
            _ = tf.cuda_numpy(xn), tf.cuda_numpy(vn)
            s = time.perf_counter() - t0
            out[key] = {"bodies": nb, "ms": ms, "ginteractions_per_s": nb * nb / ms / 1e6,
                        "roofline": {"bound": "fp32", "achieved": 20.0 * nb * nb / ms / 1e9, "peak": fp32_peak, "unit": "TFLOP/s",
                                     "frac": 20.0 * nb * nb / ms / 1e9 / fp32_peak},
                        "note": "the reference's " + what + " compiled by tf.compile on the CUDA backend (drop-in path)",
                        "verify": {"ok": bool(err <= 1e-4), "max_err_over_summed_term_magnitudes_vs_float64_on_256_bodies": err},
                        "e2e": {"value": nb * nb / s / 1e9, "unit": "Ginteractions/s", "h2d_bytes_per_step": 12 * nb, "d2h_bytes_per_step": 24 * nb}}

    m = 4096 if quick else 8192
    shared = {}

    def matrix():
        if "a" not in shared:
            shared["ha"] = rng.random((m, m), dtype=np.float32)
            shared["a"] = tf.cuda_tensor(shared["ha"])
        return shared["a"], shared["ha"]

    def reduce_section():
        # row reductions over 8192^2 fp32: one read of A
        a, ha = matrix()
        ref = {"sum": ha.sum(axis=1, dtype=np.float64), "max": ha.max(axis=1).astype(np.float64),
               "norm": np.sqrt((ha.astype(np.float64) ** 2).sum(axis=1))}
        for op in ("sum", "max", "norm"):
            ms = time_call(tf, lambda: tf.cuda_reduce(a, -1, op), 20)
            err = rel_err(tf.cuda_numpy(tf.cuda_reduce(a, -1, op)), ref[op], 1e-30)
            out[f"reduce_{op}"] = {"shape": [m, m], "ms": ms, "roofline": {"bound": "hbm", "achieved": m * m * 4 / ms / 1e6, "peak": hbm, "unit": "GB/s",
                                   "frac": m * m * 4 / ms / 1e6 / hbm}, "verify": {"ok": bool(err <= 1e-5), "max_rel_err_vs_float64": err}}
        red = workloads.compile_row_reductions(tf, m)  # the compiled program: 4 library reductions -> 4 reads of A
        ms = time_call(tf, lambda: red(a), 10)
        outs = red(a)
        errs = [rel_err(tf.cuda_numpy(outs[0]), ref["sum"], 1e-30), rel_err(tf.cuda_numpy(outs[1]), ref["max"], 1e-30),
                rel_err(tf.cuda_numpy(outs[2]), ref["sum"] / m, 1e-30), rel_err(tf.cuda_numpy(outs[3]), ref["norm"], 1e-30)]
        pin = tf.cuda_pinned_array([m, m], "float32")
        pin[...] = ha
        dev = tf.cuda_tensor(ha)
        t0 = time.perf_counter()
        tf.cuda_upload(dev, pin)
        res = [tf.cuda_numpy(o) for o in red(dev)]
        s = time.perf_counter() - t0
        out["reduce_program_4ops"] = {"shape": [m, m], "ms": ms, "roofline": {"bound": "hbm", "achieved": 4 * m * m * 4 / ms / 1e6, "peak": hbm, "unit": "GB/s",
                                      "frac": 4 * m * m * 4 / ms / 1e6 / hbm, "note": "tf.sum/max/mean/norm in one compiled program, each a library call reading A once"},
                                      "verify": {"ok": bool(max(errs) <= 1e-5), "max_rel_err_vs_float64": max(errs)},
                                      "e2e": {"value": 4 * m * m * 4 / s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": 4 * m * m, "d2h_bytes_per_step": 16 * m, "ms": s * 1e3}}
        os.environ["TFCUDA_LIBRARY"] = "0"
        try:
            red_generic = workloads.compile_row_reductions(tf, m)
        finally:
            os.environ.pop("TFCUDA_LIBRARY")
        ms = time_call(tf, lambda: red_generic(a), 5)
        out["reduce_program_4ops_generic_lowering"] = {"shape": [m, m], "ms": ms, "gbs_one_read": m * m * 4 / ms / 1e6}

    def scan_section():
        # inclusive prefix sum along the rows of the same matrix: one read + one write
        a, ha = matrix()
        ms = time_call(tf, lambda: tf.cuda_prefix_sum(a, -1), 10)
        got = tf.cuda_numpy(tf.cuda_prefix_sum(a, -1))
        rows = rng.choice(m, 64, replace=False)
        err = rel_err(got[rows], np.cumsum(ha[rows].astype(np.float64), axis=1), 1e-30)
        out["prefix_sum_rows"] = {"shape": [m, m], "ms": ms, "roofline": {"bound": "hbm", "achieved": 2 * m * m * 4 / ms / 1e6, "peak": hbm, "unit": "GB/s",
                                  "frac": 2 * m * m * 4 / ms / 1e6 / hbm}, "verify": {"ok": bool(err <= 1e-5), "max_rel_err_vs_float64_on_64_rows": err}}

    def matmul_section():
        a, ha = matrix()
        hb = rng.random((m, m), dtype=np.float32)
        b = tf.cuda_tensor(hb)
        rows = rng.choice(m, 64, replace=False)
        ref = ha[rows].astype(np.float64) @ hb.astype(np.float64)

        def err_of(c):
            return rel_err(tf.cuda_numpy(c)[rows], ref, 1e-30)

        mm_prog = workloads.compile_matmul(tf)
        ms = time_call(tf, lambda: mm_prog(a, b), 10, warm=3)
        err = err_of(mm_prog(a, b))
        out["matmul_program"] = {"shape": [m, m, m], "ms": ms, "tflops": 2.0 * m ** 3 / ms / 1e9,
                                 "note": "`a @ b` in a compiled program (library call; precision mode " + os.environ.get("TFCUDA_MATMUL_MODE", "1") +
                                         ": 0 = TF32, 1 = 3xTF32 fp32-accurate (default), 2 = FFMA)",
                                 "verify": {"ok": bool(err <= 1e-3), "max_rel_err_vs_float64_on_64_rows": err}}
        ms = time_call(tf, lambda: tf.cuda_matmul(a, b, 2), 3, warm=1)
        err = err_of(tf.cuda_matmul(a, b, 2))
        out["matmul_ffma"] = {"shape": [m, m, m], "ms": ms, "tflops": 2.0 * m ** 3 / ms / 1e9,
                              "roofline": {"bound": "fp32", "achieved": 2.0 * m ** 3 / ms / 1e9, "peak": fp32_peak, "unit": "TFLOP/s", "frac": 2.0 * m ** 3 / ms / 1e9 / fp32_peak},
                              "verify": {"ok": bool(err <= 5e-5), "max_rel_err_vs_float64_on_64_rows": err}}
        # bars at K = 8192: TF32 1e-3 (north_star's matmul bar); 3xTF32 3e-4 (the tensor core's truncating fp32 accumulation remains)
        for mode, name, mult, bar in ((0, "matmul_tcgen05_tf32", 1.0, 1e-3), (1, "matmul_tcgen05_3xtf32", 3.0, 3e-4)):
            ms = time_call(tf, lambda: tf.cuda_matmul(a, b, mode), 10, warm=3)
            err = err_of(tf.cuda_matmul(a, b, mode))
            out[name] = {"shape": [m, m, m], "ms": ms, "tflops": 2.0 * m ** 3 / ms / 1e9,
                         "roofline": {"bound": "tensor", "achieved": mult * 2.0 * m ** 3 / ms / 1e9, "peak": tf32_peak, "unit": "TFLOP/s",
                                      "frac": mult * 2.0 * m ** 3 / ms / 1e9 / tf32_peak,
                                      "note": "tensor-pipe flops (3 TF32 products per fp32 product in 3xTF32 mode) incl. the transpose/split pre-pass; "
                                              "tf32 dense peak taken as half of the measured bf16 peak"},
                         "verify": {"ok": bool(err <= bar), "bar": bar, "max_rel_err_vs_float64_on_64_rows": err}}
        pin_a, pin_b, pin_c = (tf.cuda_pinned_array([m, m], "float32") for _ in range(3))
        pin_a[...] = ha
        pin_b[...] = hb
        da, db = tf.cuda_tensor(ha), tf.cuda_tensor(hb)
        t0 = time.perf_counter()
        tf.cuda_upload(da, pin_a)
        tf.cuda_upload(db, pin_b)
        tf.cuda_download(tf.cuda_matmul(da, db, 0), pin_c)
        s = time.perf_counter() - t0
        out["matmul_tcgen05_tf32"]["e2e"] = {"value": 2.0 * m ** 3 / s / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": 8 * m * m, "d2h_bytes_per_step": 4 * m * m, "ms": s * 1e3}

    def scatter_section():
        # tf.scatterAdd / the autodiff of `load` (Implementations.cpp:185-194): dst[index[i]] += src[i], library kernel (warp-aggregated
        # red.global.add).  Algorithmic traffic: 8 B read per element (index + value); two destination sizes: L2-resident and HBM-sized
        n = 1 << (22 if quick else 26)
        for tag, dst_n in (("scatter_add_1m_bins", 1 << 20), ("scatter_add_64m_bins", 1 << 26)):
            idx = rng.integers(0, dst_n, n, dtype=np.int64).astype(np.int32)
            src = rng.random(n, dtype=np.float32)
            d_idx, d_src = tf.cuda_tensor(idx), tf.cuda_tensor(src)
            dst = tf.cuda_tensor(np.zeros(dst_n, np.float32))
            ms = time_call(tf, lambda: tf.cuda_scatter_add(dst, d_idx, d_src), 10)
            check = tf.cuda_tensor(np.zeros(dst_n, np.float32))
            tf.cuda_scatter_add(check, d_idx, d_src)
            want = np.bincount(idx, weights=src.astype(np.float64), minlength=dst_n)
            err = rel_err(tf.cuda_numpy(check), want, 1e-3 * float(want.max()))
            out[tag] = {"n": n, "bins": dst_n, "ms": ms, "gelements_per_s": n / ms / 1e6,
                        "roofline": {"bound": "hbm", "achieved": 8.0 * n / ms / 1e6, "peak": hbm, "unit": "GB/s", "frac": 8.0 * n / ms / 1e6 / hbm,
                                     "note": "8 B read per element; the atomics themselves resolve in L2 (random 4-byte RMWs: 32 B sectors)"},
                        "verify": {"ok": bool(err <= 1e-5), "max_rel_err_vs_float64_bincount": err}}

    section("radix_sort", sort_section)
    section("scatter_add", scatter_section)
    section("nbody", nbody_section)
    section("reduce", reduce_section)
    section("prefix_sum", scan_section)
    section("matmul", matmul_section)
    return out


def run_json_subprocess(cmd, timeout):
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd="/tmp")
    for line in reversed(r.stdout.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line), r
    return None, r


def cpu_baseline(args):
    """Rank 0, N=1: the reference's C++/OpenMP backend on this box's host cores, bounded sample of the same workload."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(args.cpu_steps), "--warmup", "1", "--size", str(args.size)]
    cores = host_cores()
    try:
        j, r = run_json_subprocess(cmd, 900)
        if j is None:
            return {"value": None, "unit": "GB/s", "cores": cores, "kind": "reference", "sample": "reference arm printed no JSON: " + r.stderr[-200:]}
        if "cpu_baseline" in j:
            return j["cpu_baseline"]
        return {"value": None, "unit": "GB/s", "cores": cores, "kind": "reference", "sample": j.get("unavailable", "unavailable")}
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": "GB/s", "cores": cores, "kind": "reference", "sample": f"failed: {e}"}


def cpu_extras(args):
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", "extras"]
    try:
        j, r = run_json_subprocess(cmd, 900)
        if j is None or "extras" not in j:
            return {"error": "reference extras leg printed no JSON: " + r.stderr[-200:]}
        for v in j["extras"].values():
            v["cores"] = j["cores"]
            v["kind"] = "reference"
        return j["extras"]
    except Exception as e:  # noqa: BLE001
        return {"error": f"failed: {e}"}


def nbody_approx_leg():
    """The reference's n-body programs once more, in a process whose kernels are compiled with approximate division / square root
    (`tf.initialize(tf.cuda, "--prec-div=false --prec-sqrt=false")`: 2 ulp per operation, inside north_star's 1e-5; IEEE is the default)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--workload", "nbody"]
    env = dict(os.environ, TFCUDA_KERNEL_OPTIONS="--prec-div=false --prec-sqrt=false")
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd="/tmp", env=env)
        for line in reversed(r.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return {"error": "no JSON: " + r.stderr[-200:]}
    except Exception as e:  # noqa: BLE001
        return {"error": f"failed: {e}"}


def nca_single_gpu(args):
    """The 1-GPU point of the data-parallel NCA config, in its own process (fresh pool, no interference with the fluid numbers)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--workload", "nca", "--steps", str(args.nca_iters), "--warmup", "3", "--no-single"]
    try:
        j, r = run_json_subprocess(cmd, 1500)
        if j is None:
            return {"error": "NCA leg printed no JSON: " + (r.stderr[-300:] or r.stdout[-300:])}
        keep = ("metric", "value", "unit", "n_gpus", "ms_per_step", "scaling", "config", "gpu_launches", "loss_after", "build_seconds", "graph",
                "host_issue_ms_per_step", "verify", "emitter")
        return {k: j[k] for k in keep if k in j}
    except Exception as e:  # noqa: BLE001
        return {"error": f"failed: {e}"}


def summarize_extra(extra):
    """One short row per config (the driver keeps the tail of the line: this sits at the end)."""
    rows = {}
    for k, v in extra.items():
        if not isinstance(v, dict):
            rows[k] = v
            continue
        row = {}
        for key in ("ms", "gkeys_per_s", "ginteractions_per_s", "tflops"):
            if key in v:
                row[key] = round(v[key], 4)
        if "roofline" in v:
            row["frac"] = round(v["roofline"]["frac"], 4)
            row["bound"] = v["roofline"]["bound"]
        if "verify" in v:
            row["ok"] = v["verify"].get("ok")
        if "e2e" in v:
            row["e2e"] = round(v["e2e"]["value"], 4)
        if "cpu_baseline" in v and isinstance(v["cpu_baseline"], dict) and v["cpu_baseline"].get("value") is not None:
            row["cpu"] = round(v["cpu_baseline"]["value"], 5)
        rows[k] = row
    return rows


def make_line(args, world, res):
    """The ONE JSON line of the CUDA arm (without cpu_baseline / extra / nca_dp, which main() adds at N = 1)."""
    cfg = fluid_config(args.size)  # exactly the dict the reference arm prints
    return {
        "metric": METRIC, "value": res["value"], "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": fluid_warmup(args),
        "ms_per_step": res["ms"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg, "parallelism": "replicas only (single-device program)" if world > 1 else "1 GPU",
        "bytes_per_step_counted_live": res.get("counted_bytes"), "gpu_launches": int(res["launches"]), "clocks": res["clocks"], "roofline": res["roofline"], "e2e": res["e2e"],
        "graph_replay": res.get("graph"), "verify": res.get("verify"), "emitter": res.get("emitter"), "top_kernels": res["records"],
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--cpu-steps", type=int, default=40, help="steps of the CPU baseline sample (about 10-30 s of host work)")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs (profiling runs)")
    ap.add_argument("--no-verify", action="store_true", help="skip the parity check of the timed program against the reference backend")
    ap.add_argument("--no-nca", action="store_true", help="skip the 1-GPU NCA point at N = 1")
    ap.add_argument("--no-single", action="store_true", help="NCA at N > 1: skip rank 0's single-GPU reference runs")
    ap.add_argument("--quick", action="store_true", help="smaller extra workloads (development)")
    ap.add_argument("--workload", default=None, choices=["fluid", "nca", "extras", "nbody"],
                    help="default: fluid at N = 1, nca (the config that shards) under torchrun with N > 1")
    ap.add_argument("--verify-steps", type=int, default=10)
    ap.add_argument("--dump-state", default=None, help="reference arm: save the fields after --verify-steps steps from rest")
    ap.add_argument("--nca-batch", type=int, default=256, help="GLOBAL batch (split across ranks: strong scaling)")
    ap.add_argument("--nca-weak", action="store_true", help="weak scaling: --nca-batch / --nca-pool are PER GPU (e.g. --nca-batch 32 --nca-pool 128)")
    ap.add_argument("--nca-grid", type=int, default=128)
    ap.add_argument("--nca-pool", type=int, default=1024)
    ap.add_argument("--nca-steps", type=int, default=25, help="CA steps per training iteration")
    ap.add_argument("--nca-iters", type=int, default=8, help="timed iterations of the 1-GPU NCA point in the N = 1 line")
    ap.add_argument("--nca-matmul", default="tf32", choices=["tf32", "3xtf32", "fp32"],
                    help="precision of the NCA programs' matmuls (tf.initialize option --tf-matmul): tf32 = one tensor-core product, inside "
                         "north_star's 1e-3 matmul / gradient bar (tests/test_nca_gpu.py passes in this mode); 3xtf32 = the backend's default")
    ap.add_argument("--nca-profile", action="store_true", help="add a per-kernel profile of one iteration to the NCA line")
    ap.add_argument("--nca-mono", action="store_true", help="run the reference's single program instead of the split step (1 GPU only)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload is None:
        args.workload = "fluid" if (world == 1 or args.impl == "reference") else "nca"
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "nca":
        from tensorfrost_b200 import nca_dp
        return nca_dp.bench_main(args)
    if args.workload == "nbody":  # the n-body section alone (used by nbody_approx_leg with other kernel compile options)
        import tensorfrost_b200
        with quiet_stdout():
            tf = tensorfrost_b200.load()
            res = bench_extra(tf, read_peaks(), args.quick, only=("nbody",))
        print(json.dumps(res))
        return 0

    # the 1-GPU point of the data-parallel NCA config runs first, in its own process, while this one holds no device memory
    nca_line = nca_single_gpu(args) if (world == 1 and not args.no_nca) else None
    dist, rank, world = dist_setup(args.gpus)
    import tensorfrost_b200
    with quiet_stdout():
        tf = tensorfrost_b200.load()
        peaks = read_peaks()
        res = bench_fluid(tf, dist, rank, world, args, peaks)
        extra = None
        if rank == 0 and world == 1 and not args.no_extra:
            extra = bench_extra(tf, peaks, args.quick)
    if rank == 0:
        line = make_line(args, world, res)
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args)
            if extra is not None:
                base = cpu_extras(args)
                for name, cpu in base.items():
                    for key in extra:
                        if key.startswith(name) and isinstance(extra[key], dict):
                            extra[key]["cpu_baseline"] = cpu
        if nca_line is not None:
            line["nca_dp"] = nca_line
        if extra is not None and not args.no_cpu:
            approx = nbody_approx_leg()
            for key in ("nbody_program", "nbody_loop_program"):
                if key in approx:
                    approx[key]["note"] += "; kernels compiled with --prec-div=false --prec-sqrt=false (user option of tf.initialize; IEEE is the default)"
                    extra[key + "_approx_div_sqrt"] = approx[key]
            if "error" in approx:
                extra["nbody_approx_error"] = approx["error"]
        if extra is not None:
            line["extra"] = extra
            line["extra_summary"] = summarize_extra(extra)
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    try:
        sys.exit(main())
    except BaseException as e:  # noqa: BLE001 - every rank's traceback must reach stderr (torchrun shows only the exit code)
        if not isinstance(e, SystemExit):
            import traceback
            sys.stderr.write(f"[bench rank {os.environ.get('RANK', '0')}] " + traceback.format_exc())
            sys.stderr.flush()
            sys.exit(1)
        raise
