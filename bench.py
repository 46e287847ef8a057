#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200 CUDA backend for TensorFrost programs.

Workload (BASELINE.json configs[1]): one step of the 2-D Eulerian fluid simulation on a 2048 x 2048 fp32 grid
(the reference's examples/Simulation/fluid_simulation.ipynb program, 15 fused kernels / 43 dispatches per step),
traced by the unchanged TensorFrost frontend and executed by the CUDA backend (emitter + libtfcuda.so).

metric = fused-kernel HBM GB/s = ALGORITHMIC bytes of a step / device time of a step, where the algorithmic bytes
are the sum over the step's dispatches of the size of every tensor bound to the dispatch, each counted once
(counted live by the runtime's profiler).  `extra` carries BASELINE.json's other metrics
(radix-sort Gkeys/s at 2^28 keys, n-body Ginteractions/s at 262144 bodies, row reductions and matmul at 8192^2),
each with its own roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size 2048] [--no-extra]

N > 1: the fluid step does not shard (single-device program) -> N independent replicas, one process per
GPU, barrier + max-over-ranks timing, scaling "weak".  The data-parallel NCA config is benchmarked by --workload nca.
--impl reference: the reference's own C++/OpenMP backend (oracle/_ref) on the host cores, same program and size.
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


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


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


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own C++/OpenMP backend on the host cores
# ------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref, "TensorFrost")):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (reference module built by oracle/build_ref.sh) is missing"}))
        return 0
    sys.path.insert(0, ref)
    import numpy as np
    n = args.size
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)  # the reference prints import / compile chatter on stdout; keep stdout for the ONE JSON line
    try:
        import TensorFrost as tf
        tf.initialize(tf.cpu)  # the reference's default flags: -O3 -ffast-math -fopenmp
        from tensorfrost_b200 import workloads
        fluid = workloads.load_fluid(tf, n, n)
        state = [tf.tensor(a) for a in workloads.fluid_inputs(n, n)]
        for _ in range(args.warmup):
            state, _ = workloads.fluid_step(fluid, state)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            state, _ = workloads.fluid_step(fluid, state)
        dt = time.perf_counter() - t0
    finally:
        os.dup2(saved, 1)
    step_bytes = fluid_step_bytes_nominal(n)
    value = step_bytes * args.steps / dt / 1e9
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": "fused-kernel HBM GB/s", "value": value, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"fluid_simulation {n}x{n} fp32, 1 step = 43 dispatches (reference C++/OpenMP backend, host cores)",
                   "bytes_per_step": step_bytes},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores, "kind": "reference",
                         "sample": f"{args.steps} steps of the {n}x{n} fluid program on tf.cpu (-O3 -ffast-math -fopenmp)"},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def fluid_step_bytes_nominal(n):
    """Algorithmic bytes of one fluid step, scaled from the survey's count at 2048^2 (817 MB).  The CUDA arm
    counts the same quantity live (every tensor bound to every dispatch, once); this closed form serves the CPU arm, which
    has no dispatch profiler, so both arms divide the SAME byte count by their own time."""
    return 817.0e6 * (n * n) / (2048.0 * 2048.0)


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


def ncu_traffic(kernel_name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of an emitted fluid kernel, from the committed `ncu --set full` capture
    of this same command (profiles/README.md); None when the capture does not hold that kernel."""
    import csv
    path = os.path.join(ROOT, "profiles", "r01b_fluid_ncu_full_summary.csv")
    try:
        rows = list(csv.reader(open(path)))
        col = {h: i for i, h in enumerate(rows[0])}
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        vals = []
        for r in rows[2:]:
            if r[col["Kernel Name"]] == kernel_name:
                rd = float(r[col["dram__bytes_read.sum"]]) * unit[rows[1][col["dram__bytes_read.sum"]]]
                wr = float(r[col["dram__bytes_write.sum"]]) * unit[rows[1][col["dram__bytes_write.sum"]]]
                vals.append(rd + wr)
        return sum(vals) / len(vals) if vals else None
    except (OSError, KeyError, ValueError, IndexError):
        return None


def ncu_limiter(kernel_name):
    """What the same capture says limits that kernel: issue-slot, DRAM and L1 utilisation (percent of peak)."""
    import csv
    path = os.path.join(ROOT, "profiles", "r01b_fluid_ncu_full_summary.csv")
    try:
        rows = list(csv.reader(open(path)))
        col = {h: i for i, h in enumerate(rows[0])}
        for r in rows[2:]:
            if r[col["Kernel Name"]] == kernel_name:
                return {"issue_active_pct": float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
                        "dram_pct": float(r[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
                        "l1tex_pct": float(r[col["l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]]),
                        "warp_instructions": float(r[col["smsp__inst_executed.sum"]])}
    except (OSError, KeyError, ValueError, IndexError):
        pass
    return None


def dominant(records, steps):
    """Pick the kernel with the largest share of device time; return its roofline fields."""
    total = sum(r["total_ms"] for r in records) or 1.0
    top = max(records, key=lambda r: r["total_ms"])
    per_launch_ms = top["total_ms"] / max(top["launches"], 1)
    per_launch_bytes = top["bytes"] / max(top["launches"], 1)
    return top, per_launch_ms, per_launch_bytes, top["total_ms"] / total


def bench_fluid(tf, dist, rank, world, args, peaks):
    import numpy as np
    from tensorfrost_b200 import workloads
    n = args.size
    fluid = workloads.load_fluid(tf, n, n)
    host_inputs = workloads.fluid_inputs(n, n)
    state = [tf.cuda_tensor(a) for a in host_inputs]
    for _ in range(max(args.warmup, 3)):
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
    # ---- per-kernel profile of the same K steps (event pair per launch) -> bytes per step and the dominant kernel ----
    tf.cuda_profile_reset()
    tf.cuda_profile_enable(True)
    for _ in range(args.steps):
        state, _ = workloads.fluid_step(fluid, state)
    tf.cuda_profile_enable(False)
    records = tf.cuda_profile_records()
    step_bytes = sum(r["bytes"] for r in records) / args.steps
    top, top_ms, top_bytes, share = dominant(records, args.steps)
    value = world * step_bytes * args.steps / (ms / 1e3) / 1e9
    achieved = top_bytes / (top_ms / 1e3) / 1e9 if top_ms > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": ncu_traffic(top["name"]) if n == 2048 else None, "traffic_source": "profiles/r01b_fluid_ncu_full_summary.csv (ncu --set full, same command)",
                "kernel": top["name"], "ncu": ncu_limiter(top["name"]) if n == 2048 else None, "share_of_step": share, "launch_ms": top_ms, "algorithmic_bytes_per_launch": top_bytes,
                "peak_source": peaks["source"],
                "whole_step": {"achieved": step_bytes / (sum(r["total_ms"] for r in records) / args.steps / 1e3) / 1e9,
                               "note": "sum of algorithmic bytes / sum of kernel times over all 43 dispatches"}}
    # ---- e2e: the same step through the public API with HOST buffers: pinned H2D of the 4 fields + D2H of the 4 result fields ----
    pinned = [tf.cuda_pinned_array([n, n], "float32") for _ in range(4)]
    for p, a in zip(pinned, host_inputs[:4]):
        p[...] = a
    e2e_state = [tf.cuda_tensor(a) for a in host_inputs]
    direct = hasattr(tf, "cuda_download")
    if direct:
        try:
            tf.cuda_download(e2e_state[0], pinned[0])
            pinned[0][...] = host_inputs[0]
        except Exception:  # noqa: BLE001 - older module: stage through tf.cuda_numpy
            direct = False
    barrier(dist, tf)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for k in range(4):
            tf.cuda_upload(e2e_state[k], pinned[k])
        e2e_state, _ = workloads.fluid_step(fluid, e2e_state)
        if direct:
            for k in range(4):
                tf.cuda_download(e2e_state[k], pinned[k])  # device -> the same page-locked arrays, no pageable staging copy
        else:
            outs = [tf.cuda_numpy(e2e_state[k]) for k in range(4)]
            for p, o in zip(pinned, outs):
                p[...] = o
    tf.cuda_synchronize()
    e2e_s = max_over_ranks(dist, time.perf_counter() - t0)
    e2e = {"value": world * step_bytes * args.steps / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": 4 * n * n * 4,
           "d2h_bytes_per_step": 4 * n * n * 4, "ms_per_step": e2e_s / args.steps * 1e3,
           "path": "pinned host arrays <-> device, tf.cuda_upload / " + ("tf.cuda_download" if direct else "tf.cuda_numpy + host copy")}
    return {"value": value, "ms": ms, "launches": launches, "clocks": clocks, "roofline": roofline, "e2e": e2e, "step_bytes": step_bytes,
            "records": sorted(records, key=lambda r: -r["total_ms"])[:16]}


def time_call(tf, fn, iters, warm=3):
    for _ in range(warm):
        fn()
    tf.cuda_synchronize()
    tf.cuda_timer_begin()
    for _ in range(iters):
        fn()
    return tf.cuda_timer_end() / iters


def bench_extra(tf, peaks, quick):
    """BASELINE.json's other metrics at the configs' sizes, each against its own roofline.  Every section is independent: a failure is
    recorded under its name and does not take the headline line down with it."""
    import numpy as np
    from tensorfrost_b200 import workloads
    out = {}
    rng = np.random.default_rng(0)
    hbm = peaks["hbm_gbs"]
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    tf32_peak = peaks["bf16_tflops"] / 2

    def section(name, fn):
        try:
            fn()
        except Exception as e:  # noqa: BLE001
            out[name + "_error"] = f"{type(e).__name__}: {e}"[:300]

    def sort_section():
        # radix sort, 2^28 uint32 keys (keys-only: 36 B/key algorithmic; key+value: 68 B/pair)
        n = 1 << (24 if quick else 28)
        keys = tf.cuda_tensor(rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32))
        ms = time_call(tf, lambda: tf.cuda_radix_sort(keys), 5)
        out["radix_sort_keys"] = {"n": n, "ms": ms, "gkeys_per_s": n / ms / 1e6, "roofline": {"bound": "hbm", "achieved": 36.0 * n / ms / 1e6, "peak": hbm,
                                  "unit": "GB/s", "frac": 36.0 * n / ms / 1e6 / hbm}}
        vals = tf.cuda_tensor(np.arange(n, dtype=np.uint32))
        ms = time_call(tf, lambda: tf.cuda_radix_sort(keys, vals), 5)
        out["radix_sort_pairs"] = {"n": n, "ms": ms, "gkeys_per_s": n / ms / 1e6, "roofline": {"bound": "hbm", "achieved": 68.0 * n / ms / 1e6, "peak": hbm,
                                   "unit": "GB/s", "frac": 68.0 * n / ms / 1e6 / hbm}}
        sort_prog = workloads.compile_sort(tf, with_values=True)  # tf.sort.radix inside a compiled program
        ms = time_call(tf, lambda: sort_prog(keys, vals), 5)
        out["radix_sort_pairs_program"] = {"n": n, "ms": ms, "gkeys_per_s": n / ms / 1e6, "note": "tf.sort.radix(keys, values) traced by tf.compile: one library call + output copies"}

    def nbody_section():
        # n-body, 262144 bodies: library kernel and the generic emitter on the reference program
        nb = 32768 if quick else 262144
        x = tf.cuda_tensor((5.0 * rng.standard_normal((nb, 3))).astype(np.float32))
        v = tf.cuda_tensor(np.zeros((nb, 3), np.float32))
        ms = time_call(tf, lambda: tf.cuda_nbody_step(x, v), 3, warm=1)
        out["nbody_library"] = {"bodies": nb, "ms": ms, "ginteractions_per_s": nb * nb / ms / 1e6,
                                "roofline": {"bound": "fp32", "achieved": 20.0 * nb * nb / ms / 1e9, "peak": fp32_peak, "unit": "TFLOP/s",
                                             "frac": 20.0 * nb * nb / ms / 1e9 / fp32_peak, "note": "20 flop/interaction (SURVEY \u00a78d C3); CUDA-core bound, not HBM"}}
        nbody = workloads.compile_nbody(tf)
        ms = time_call(tf, lambda: nbody(x, v), 2, warm=1)
        out["nbody_emitted"] = {"bodies": nb, "ms": ms, "ginteractions_per_s": nb * nb / ms / 1e6}

    m = 4096 if quick else 8192
    shared = {}

    def reduce_section():
        # row reductions over 8192^2 fp32: one read of A
        a = shared["a"] = tf.cuda_tensor(rng.random((m, m), dtype=np.float32))
        for op in ("sum", "max", "norm"):
            ms = time_call(tf, lambda: tf.cuda_reduce(a, -1, op), 20)
            out[f"reduce_{op}"] = {"shape": [m, m], "ms": ms, "roofline": {"bound": "hbm", "achieved": m * m * 4 / ms / 1e6, "peak": hbm, "unit": "GB/s",
                                   "frac": m * m * 4 / ms / 1e6 / hbm}}
        red = workloads.compile_row_reductions(tf, m)  # the compiled program: 4 library reductions -> 4 reads of A
        ms = time_call(tf, lambda: red(a), 10)
        out["reduce_program_4ops"] = {"shape": [m, m], "ms": ms, "roofline": {"bound": "hbm", "achieved": 4 * m * m * 4 / ms / 1e6, "peak": hbm, "unit": "GB/s",
                                      "frac": 4 * m * m * 4 / ms / 1e6 / hbm, "note": "tf.sum/max/mean/norm in one compiled program, each a library call reading A once"}}
        os.environ["TFCUDA_LIBRARY"] = "0"
        try:
            red_generic = workloads.compile_row_reductions(tf, m)
        finally:
            os.environ.pop("TFCUDA_LIBRARY")
        ms = time_call(tf, lambda: red_generic(a), 5)
        out["reduce_program_4ops_generic_lowering"] = {"shape": [m, m], "ms": ms, "gbs_one_read": m * m * 4 / ms / 1e6}

    def scan_section():
        # inclusive prefix sum along the rows of the same matrix: one read + one write
        a = shared.get("a")
        if a is None:
            a = shared["a"] = tf.cuda_tensor(rng.random((m, m), dtype=np.float32))
        ms = time_call(tf, lambda: tf.cuda_prefix_sum(a, -1), 10)
        out["prefix_sum_rows"] = {"shape": [m, m], "ms": ms, "roofline": {"bound": "hbm", "achieved": 2 * m * m * 4 / ms / 1e6, "peak": hbm, "unit": "GB/s",
                                  "frac": 2 * m * m * 4 / ms / 1e6 / hbm}}

    def matmul_section():
        a = shared.get("a")
        if a is None:
            a = shared["a"] = tf.cuda_tensor(rng.random((m, m), dtype=np.float32))
        b = tf.cuda_tensor(rng.random((m, m), dtype=np.float32))
        mm_prog = workloads.compile_matmul(tf)
        ms = time_call(tf, lambda: mm_prog(a, b), 10, warm=3)
        out["matmul_program"] = {"shape": [m, m, m], "ms": ms, "tflops": 2.0 * m ** 3 / ms / 1e9, "note": "`a @ b` in a compiled program (library call, 3xTF32 mode by default)"}
        ms = time_call(tf, lambda: tf.cuda_matmul(a, b, 2), 3, warm=1)
        out["matmul_ffma"] = {"shape": [m, m, m], "ms": ms, "tflops": 2.0 * m ** 3 / ms / 1e9,
                              "roofline": {"bound": "fp32", "achieved": 2.0 * m ** 3 / ms / 1e9, "peak": fp32_peak, "unit": "TFLOP/s", "frac": 2.0 * m ** 3 / ms / 1e9 / fp32_peak}}
        for mode, name, mult in ((0, "matmul_tcgen05_tf32", 1.0), (1, "matmul_tcgen05_3xtf32", 3.0)):
            ms = time_call(tf, lambda: tf.cuda_matmul(a, b, mode), 10, warm=3)
            out[name] = {"shape": [m, m, m], "ms": ms, "tflops": 2.0 * m ** 3 / ms / 1e9,
                         "roofline": {"bound": "tensor", "achieved": mult * 2.0 * m ** 3 / ms / 1e9, "peak": tf32_peak, "unit": "TFLOP/s",
                                      "frac": mult * 2.0 * m ** 3 / ms / 1e9 / tf32_peak,
                                      "note": "tensor-pipe flops (3 TF32 products per fp32 product in 3xTF32 mode) incl. the transpose/split pre-pass; "
                                              "tf32 dense peak taken as half of the measured bf16 peak"}}

    section("radix_sort", sort_section)
    section("nbody", nbody_section)
    section("reduce", reduce_section)
    section("prefix_sum", scan_section)
    section("matmul", matmul_section)
    return out


def cpu_baseline(args):
    """Rank 0, N=1: the reference's C++/OpenMP backend on this box's host cores, bounded sample of the same workload."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(args.cpu_steps), "--warmup", "1", "--size", str(args.size)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd="/tmp")
        for line in reversed(r.stdout.strip().splitlines()):
            if line.startswith("{"):
                j = json.loads(line)
                if "cpu_baseline" in j:
                    return j["cpu_baseline"]
                return {"value": None, "unit": "GB/s", "cores": os.cpu_count(), "kind": "reference", "sample": j.get("unavailable", "unavailable")}
        return {"value": None, "unit": "GB/s", "cores": os.cpu_count(), "kind": "reference", "sample": "reference arm printed no JSON: " + r.stderr[-200:]}
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": "GB/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {e}"}


def make_line(args, world, res):
    """The ONE JSON line of the CUDA arm (without cpu_baseline / extra, which main() adds at N = 1)."""
    return {
        "metric": "fused-kernel HBM GB/s", "value": res["value"], "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": res["ms"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"fluid_simulation {args.size}x{args.size} fp32 (BASELINE configs[1]), 1 step = 43 dispatches of 15 emitted kernels",
                   "bytes_per_step": res["step_bytes"], "l2": "per-step working set (~20 fields x 16.8 MB) exceeds the 126 MB L2; no explicit flush",
                   "parallelism": "replicas only (single-device program)" if world > 1 else "1 GPU"},
        "gpu_launches": int(res["launches"]), "clocks": res["clocks"], "roofline": res["roofline"], "e2e": res["e2e"],
        "top_kernels": res["records"],
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--quick", action="store_true", help="smaller extra workloads (development)")
    ap.add_argument("--workload", default="fluid", choices=["fluid", "nca"])
    ap.add_argument("--nca-batch", type=int, default=256, help="GLOBAL batch (split across ranks: strong scaling)")
    ap.add_argument("--nca-weak", action="store_true", help="weak scaling: --nca-batch / --nca-pool are PER GPU (e.g. --nca-batch 32 --nca-pool 128)")
    ap.add_argument("--nca-grid", type=int, default=128)
    ap.add_argument("--nca-pool", type=int, default=1024)
    ap.add_argument("--nca-steps", type=int, default=25, help="CA steps per training iteration")
    ap.add_argument("--nca-profile", action="store_true", help="add a per-kernel profile of one iteration to the NCA line")
    ap.add_argument("--nca-mono", action="store_true", help="run the reference's single program instead of the split step (1 GPU only)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "nca":
        from tensorfrost_b200 import nca_dp
        return nca_dp.bench_main(args)

    dist, rank, world = dist_setup(args.gpus)
    import tensorfrost_b200
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)  # TensorProgram prints its properties on every compile; keep stdout for the JSON line
    try:
        tf = tensorfrost_b200.load()
        peaks = read_peaks()
        res = bench_fluid(tf, dist, rank, world, args, peaks)
        extra = None
        if rank == 0 and world == 1 and not args.no_extra:
            extra = bench_extra(tf, peaks, args.quick)
    finally:
        os.dup2(saved, 1)
    if rank == 0:
        line = make_line(args, world, res)
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args)
        if extra is not None:
            line["extra"] = extra
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
