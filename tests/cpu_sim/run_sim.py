"""TEST INFRASTRUCTURE ONLY.  Executes the CUDA C++ the emitter produces — the exact text NVRTC compiles for sm_100a — on the HOST:
every case of tests/cases.py is traced by the CUDA-enabled module in codegen mode (generic lowering: library calls have no text), its
emitted kernels are compiled by g++ through tests/cpu_sim/cuda_host_shim.h, the host program the reference generated is compiled next to
tests/cpu_sim/sim_runtime.inc, and the program is run on the case's seeded inputs.  The outputs go to an .npz for the caller to compare
with the reference's golden fixtures.  What this checks without a GPU: the kernel wrapper and argument block, binding / variable order,
block and thread index arithmetic, the prelude's helper semantics, atomics, host-side loops and read/write callbacks.  What it cannot
check: anything that depends on the hardware (CUDA math library rounding, real parallel execution, barriers).

usage: python tests/cpu_sim/run_sim.py <out.npz> <case[:size[:seed]]|fluid> [...]      (own process: the backend is a process singleton)
"""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
sys.path.insert(0, TESTS)
sys.path.insert(0, ROOT)


class SimTensor(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint32)), ("dim", C.c_size_t), ("shape", C.c_size_t * 8), ("type", C.c_int)]


TYPE_OF = {"f": 0, "u": 1, "i": 2, "b": 3}


def to_words(a):
    a = np.asarray(a)
    if a.dtype == np.bool_:
        return np.ascontiguousarray(a.astype(np.uint32)), 3
    if a.dtype.kind == "f":
        return np.ascontiguousarray(a.astype(np.float32)).view(np.uint32), 0
    if a.dtype.kind == "u":
        return np.ascontiguousarray(a.astype(np.uint32)), 1
    if a.dtype.kind == "i":
        return np.ascontiguousarray(a.astype(np.int32)).view(np.uint32), 2
    raise TypeError(f"unsupported input dtype {a.dtype}")


def from_words(words, shape, type_id):
    words = words.reshape(shape)
    if type_id == 0:
        return words.view(np.float32)
    if type_id == 2:
        return words.view(np.int32)
    if type_id == 3:
        return words != 0
    return words


def coarsened(kernels):
    """[kernels whose threads carry several lanes, of which with a lane-by-lane path for blocks on the edge of the dispatch, kernels]"""
    texts = [k[0][1] + k[0][2] for k in kernels]
    lanes = [t for t in texts if "lanes per thread" in t]
    return np.array([len(lanes), sum("tf_lane" in t for t in lanes), len(texts)], np.int64)


def kernel_units(kernels, host_code):
    """C++ text of the kernels translation unit: shim + prelude + emitted kernels + one serial launcher per kernel."""
    groups = {}
    for line in host_code.splitlines():
        m = re.search(r"tf\.dispatch\((\d+),.*\{([^{}]*)\}\);\s*$", line)
        if m:
            g = [int(x) for x in m.group(2).replace(" ", "").split(",") if x]
            groups[int(m.group(1))] = (g + [1, 1, 1])[:3]
    text = ['#include "cuda_host_shim.h"', open(os.path.join(ROOT, "tensorfrost_b200", "csrc", "prelude.cuh")).read()]
    cases_ = []
    for k in kernels:
        src = k[0][1] + k[0][2]
        if "tfcuda_lib:" in src:
            continue  # a library call: no text; handled by the library callback (library_calls below)
        needs_barrier = "tf_group_barrier" in src
        m = re.search(r"void (?:__launch_bounds__\(\d+\) )?kernel_(\d+)\(", src)
        kid = int(m.group(1))
        nm = re.search(r"uint\* mem\[(\d+)\];", src)
        nv = re.search(r"uint var\[(\d+)\];", src)
        n_mem, n_var = (int(nm.group(1)) if nm else 0), int(nv.group(1))
        gx, gy, gz = groups.get(kid, (None, None, None))
        if gx is None:
            continue  # never dispatched by this program
        lb = re.search(r"// tfcuda_block: (\d+) (\d+) (\d+)", src)
        if lb:  # the block the kernel is LAUNCHED with (smaller than the IR's for coarsened kernels: each thread carries several lanes)
            gx, gy, gz = (int(x) for x in lb.groups())
        text.append(src)
        fill_mem = f"for (size_t i = 0; i < {n_mem}; i++) a.mem[i] = mem[i];" if n_mem else ""
        if needs_barrier:
            body = f"""
    blockIdx = sim_dim3{{(unsigned)b, 0, 0}};
    std::barrier<> bar({gx} * {gy} * {gz});
    sim_block_barrier = &bar;
    std::vector<std::thread> threads;
    for (unsigned z = 0; z < {gz}; z++) for (unsigned y = 0; y < {gy}; y++) for (unsigned x = 0; x < {gx}; x++)
      threads.emplace_back([&a, &bar, b, x, y, z, wgc]() {{
        gridDim = sim_dim3{{(unsigned)wgc, 1, 1}};
        blockDim = sim_dim3{{{gx}, {gy}, {gz}}};
        blockIdx = sim_dim3{{(unsigned)b, 0, 0}};
        threadIdx = sim_dim3{{x, y, z}};
        kernel_{kid}(a);
        bar.arrive_and_drop();  // an exited CUDA thread no longer takes part in the block's barriers
      }});
    for (auto& t : threads) t.join();
    sim_block_barrier = nullptr;"""
        else:
            body = f"""
    blockIdx = sim_dim3{{(unsigned)b, 0, 0}};
    for (unsigned z = 0; z < {gz}; z++) for (unsigned y = 0; y < {gy}; y++) for (unsigned x = 0; x < {gx}; x++) {{
      threadIdx = sim_dim3{{x, y, z}};
      kernel_{kid}(a);
    }}"""
        text.append(f"""
static int launch_{kid}(uint32_t** mem, size_t n_mem, const uint32_t* vars, size_t n_var, size_t wgc) {{
  if (n_mem != {n_mem} || n_var != {n_var}) return 2;
  kernel_{kid}_args a;
  {fill_mem}
  for (size_t i = 0; i < {n_var}; i++) a.var[i] = vars[i];
  gridDim = sim_dim3{{(unsigned)wgc, 1, 1}};
  blockDim = sim_dim3{{{gx}, {gy}, {gz}}};
  for (size_t b = 0; b < wgc; b++) {{{body}
  }}
  return 0;
}}""")
        cases_.append(f"    case {kid}: return launch_{kid}(mem, n_mem, vars, n_var, wgc);")
    text.append('extern "C" int sim_kernel_launch(size_t kernel_id, uint32_t** mem, size_t n_mem, const uint32_t* vars, size_t n_var, size_t wgc) {\n'
                "  switch (kernel_id) {\n" + "\n".join(cases_) + "\n    default: return 1;\n  }\n}\n")
    return "\n".join(text)


def library_calls(kernels):
    """kernel id -> (op, params, input bindings, output bindings) from the marker comments of library-call kernels."""
    calls = {}
    for k in kernels:
        m = re.search(r"// kernel_(\d+): tfcuda_lib:([a-z_]+)((?::-?\d+)*) inputs=\[([\d,]*)\] outputs=\[([\d,]*)\]", k[0][2])
        if m:
            ints = lambda t: [int(x) for x in t.split(",") if x]  # noqa: E731
            calls[int(m.group(1))] = (m.group(2), [int(x) for x in m.group(3).split(":") if x], ints(m.group(4)), ints(m.group(5)))
    return calls


REDUCE_OPS = {0: "sum", 1: "max", 2: "min", 3: "mean", 4: "norm"}  # TFCUDA_RED_* of include/tfcuda.h


def make_library_callback(calls):
    """The host stand-in for libtfcuda's library kernels: the numpy restatement of the reference algorithms (oracle/tf_oracle.py), applied
    with the same role / extent / axis conventions as DispatchCudaLibraryCall (overlay/Backend/Backends/CUDA/CudaLibrary.cpp)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tf_oracle
    dtypes = {0: np.float32, 1: np.uint32, 2: np.int32, 3: np.uint32}

    def view(t):
        shape = tuple(t.shape[d] for d in range(t.dim))
        n = int(np.prod(shape)) if shape else 1
        return np.ctypeslib.as_array(t.data, shape=(n,)).view(dtypes[t.type]).reshape(shape)

    def callback(kernel_id, n_tensors, tensors):
        if kernel_id not in calls:
            return 0
        try:
            op, params, ins, outs = calls[kernel_id]
            t = [view(tensors[i]) for i in range(n_tensors)]
            if op == "matmul":
                a, b, c = t[ins[0]], t[ins[1]], t[outs[0]]
                c[...] = tf_oracle.matmul(a, b).reshape(c.shape)
            elif op == "matmul_tn":
                a, b, c = t[ins[0]], t[ins[1]], t[outs[0]]
                m, n = a.shape[-1], b.shape[-1]
                c[...] = (a.reshape(-1, m).astype(np.float64).T @ b.reshape(-1, n).astype(np.float64)).astype(np.float32).reshape(c.shape)
            elif op == "reduce":
                src, dst = t[ins[0]], t[outs[0]]
                axis = src.ndim - 1 - params[1]  # IR dims are innermost-first
                dst[...] = np.asarray(tf_oracle.reduce(src, axis, REDUCE_OPS[params[0]])).reshape(dst.shape)
            elif op == "scan":
                src, dst = t[ins[0]], t[outs[0]]
                dst[...] = tf_oracle.prefix_sum(src, src.ndim - 1 - params[0]).reshape(dst.shape)
            elif op == "sort":
                keys = t[ins[0]]
                values = t[ins[1]] if params[0] else None
                k, v = tf_oracle.radix_sort(keys.reshape(-1), None if values is None else values.reshape(-1), max_bits=params[1])
                t[outs[0]][...] = k.reshape(t[outs[0]].shape)
                if values is not None:
                    t[outs[1]][...] = v.reshape(t[outs[1]].shape)
            else:
                return -1
            return 1
        except Exception as e:  # noqa: BLE001
            print(f"[run_sim] library call {calls.get(kernel_id)} failed: {type(e).__name__}: {e}", file=sys.stderr, flush=True)
            return -1
    return callback


LIBRARY_CB = C.CFUNCTYPE(C.c_int, C.c_size_t, C.c_size_t, C.POINTER(SimTensor))


def build(host_code, kernels, tag):
    """g++ the two translation units (emitted kernels through the shim; generated host program + sim runtime) into one library."""
    work = tempfile.mkdtemp(prefix=f"tfsim_{tag}_")
    with open(os.path.join(work, "kernels.cpp"), "w") as f:
        f.write(kernel_units(kernels, host_code))
    with open(os.path.join(work, "host.cpp"), "w") as f:
        f.write("#define main tf_program_main\n" + host_code + "\n" + open(os.path.join(HERE, "sim_runtime.inc")).read())
    so = os.path.join(work, "sim.so")
    # two translation units: the kernels need C++20 (<barrier>), the generated host program defines its own lerp() and must stay C++17
    common = ["g++", "-O2", "-pthread", "-w", "-fPIC", "-include", "math.h", "-I", HERE, "-c"]
    for std, unit in (("-std=c++20", "kernels"), ("-std=c++17", "host")):
        r = subprocess.run(common + [std, os.path.join(work, unit + ".cpp"), "-o", os.path.join(work, unit + ".o")], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed for {tag} ({unit}):\n{r.stderr[-3000:]}")
    r = subprocess.run(["g++", "-shared", "-pthread", os.path.join(work, "kernels.o"), os.path.join(work, "host.o"), "-o", so], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed for {tag}:\n{r.stderr[-3000:]}")
    lib = C.CDLL(so)
    lib.sim_run.argtypes = [C.POINTER(SimTensor), C.c_int, C.POINTER(SimTensor), C.c_int]
    lib.sim_run.restype = C.c_int
    lib.sim_error.restype = C.c_char_p
    lib.sim_free.argtypes = [C.c_void_p]
    calls = library_calls(kernels)
    if calls:
        lib._library_cb = LIBRARY_CB(make_library_callback(calls))  # keep the thunk alive with the library
        lib.sim_set_library_callback.argtypes = [LIBRARY_CB]
        lib.sim_set_library_callback(lib._library_cb)
    return lib


def run(lib, inputs, n_out, tag):
    ins = (SimTensor * max(len(inputs), 1))()
    keep = []
    for i, a in enumerate(inputs):
        words, type_id = to_words(a)
        keep.append(words)
        ins[i].data = words.ctypes.data_as(C.POINTER(C.c_uint32))
        ins[i].dim = words.ndim
        for d, s in enumerate(words.shape):
            ins[i].shape[d] = s
        ins[i].type = type_id
    outs = (SimTensor * max(n_out, 1))()
    if lib.sim_run(ins, len(inputs), outs, n_out) != 0:
        raise RuntimeError(f"sim_run failed for {tag}: {lib.sim_error().decode(errors='replace')}")
    result = []
    for i in range(n_out):
        shape = tuple(outs[i].shape[d] for d in range(outs[i].dim))
        n = int(np.prod(shape)) if shape else 1
        words = np.ctypeslib.as_array(outs[i].data, shape=(n,)).copy()
        lib.sim_free(outs[i].data)
        result.append(from_words(words, shape, outs[i].type))
    return result


class HostTensor:
    """What a program call returns in the fluid scenario: the `.numpy` the workload helpers read."""

    def __init__(self, array):
        self.numpy = array


def main():
    out_path, specs = sys.argv[1], sys.argv[2:]
    import tensorfrost_b200
    from tensorfrost_b200 import workloads
    tf = tensorfrost_b200.import_module()
    saved = os.dup(1)
    os.dup2(os.open(os.devnull, os.O_WRONLY), 1)
    result = {}
    try:
        tf.initialize(tf.codegen, "", tf.cuda_lang)
        import cases
        captured = []
        keep = []  # programs must stay alive: the module's kernel registry holds raw pointers into them (Backend/KernelManager.cpp:4-8)
        real_compile = tf.compile

        def capturing_compile(fn):
            p = real_compile(fn)
            captured.append(p)
            keep.append(p)
            return p
        tf.compile = capturing_compile
        seen = 0
        for spec in specs:
            parts = spec.split(":")
            # a trailing ":library" traces the case with the library lowerings on (marker kernels computed by the library callback)
            os.environ["TFCUDA_LIBRARY"] = "1" if parts[-1] == "library" else "0"
            if parts[-1] == "library":
                parts = parts[:-1]
            name = parts[0]
            del captured[:]
            if name == "fluid":
                n, m, steps = (int(v) for v in parts[1:4])
                fluid = workloads.load_fluid(tf, n, m)
                kernels = tf.get_all_generated_kernels()[seen:]
                seen += len(kernels)
                lib = build(fluid.compiled_code(), kernels, "fluid")
                result[f"{spec}/coarsened"] = coarsened(kernels)
                result[f"{spec}/lane_loops"] = np.array(sum("tf_lane" in k[0][2] and "lanes per thread" not in k[0][2] for k in kernels))

                def call(*state):
                    arrays = [t.numpy if isinstance(t, HostTensor) else t for t in state]
                    return [HostTensor(o) for o in run(lib, arrays, 7, "fluid")]
                state = workloads.fluid_inputs(n, m)
                state[5] = np.array([1.0, 0.0, 1.0, 1.0, 0.999, 0.999], np.float32)  # as tensorfrost_b200.workloads.fluid_parity_run
                div = canvas = None
                for step in range(steps):
                    state[4] = workloads.fluid_parity_mouse(step, n, m)
                    state, (canvas, div, _res) = workloads.fluid_step(call, state)
                outs = [t.numpy if isinstance(t, HostTensor) else t for t in state[:4]] + [div.numpy, canvas.numpy]
                for k, o in enumerate(outs):
                    result[f"{spec}/{k}"] = o
                continue
            if name == "nca":
                # the grad program of the data-parallel NCA step at the golden configuration, inputs as NcaTrainer.__init__ makes them
                from tensorfrost_b200 import nca_dp
                g = np.load(os.path.join(TESTS, "golden", "nca_step.npz"))
                batch, grid, pool_size, steps = int(g["global_batch"]), int(g["grid"]), int(g["pool_size"]), int(g["train_steps"])
                nca = workloads.load_nca(tf, batch, grid, pool_size=pool_size, train_steps=steps, channel_n=12)
                grad_step, apply_step, _mono, shapes = nca_dp.build_programs(tf, nca, steps)
                program = tf.compile(grad_step)
                kernels = tf.get_all_generated_kernels()[seen:]
                seen += len(kernels)
                apply_program = tf.compile(apply_step)
                apply_kernels = tf.get_all_generated_kernels()[seen:]
                seen += len(apply_kernels)
                rng = np.random.default_rng(0)
                hidden = 128
                fc1 = (rng.standard_normal((48, hidden)) * np.sqrt(2.0 / 48)).astype(np.float32)
                zeros = lambda *sh: np.zeros(sh, np.float32)  # noqa: E731
                trainable = [fc1, zeros(hidden), zeros(hidden, 12), zeros(12)]
                # optimizer module parameters in the order both traced programs declare them (their check_tensor lines):
                # fc1, fc1_bias, fc2, fc2_bias, filters, seed, Adam t, m[4], v[4]  (t, m, v start at zero: optimizers.py:55-63)
                opt = trainable + [workloads.nca_filters(), np.array([0], np.uint32), zeros(1)] + [np.zeros_like(t) for t in trainable] \
                    + [np.zeros_like(t) for t in trainable]
                pool, image = workloads.nca_pool(pool_size, grid, 12), workloads.nca_target(grid, 0)
                ids, fire, lr = np.asarray(g["ids"], np.int32), np.array([float(nca.CELL_FIRE_RATE)], np.float32), np.array([float(g["lr"])], np.float32)
                print(f"[run_sim] nca: grad program {len(kernels)} kernels, apply program {len(apply_kernels)} kernels", file=sys.stderr, flush=True)
                grad_lib, apply_lib = build(program.compiled_code(), kernels, "nca_grad"), build(apply_program.compiled_code(), apply_kernels, "nca_apply")
                result[f"{spec}/coarsened"] = coarsened(kernels)
                losses = []
                first = None
                for _ in range(len(g["split_losses"])):  # NcaTrainer.step: grad program -> (exchange) -> apply program
                    pool, seed, flat, state = run(grad_lib, opt + [pool, image, ids, fire], 4, "nca_grad")
                    opt[5] = seed
                    if first is None:
                        first = (pool, seed, flat, state)
                    losses.append(float(flat[-1]))
                    opt = run(apply_lib, opt + [flat, lr], 15, "nca_apply")
                for k, o in enumerate(first):
                    result[f"{spec}/{k}"] = o
                result[f"{spec}/losses"] = np.array(losses, np.float64)
                continue
            size = int(parts[1]) if len(parts) > 1 and parts[1] else None
            seed = int(parts[2]) if len(parts) > 2 else 0
            c = cases.CASES[name]
            inputs = c.make_inputs(np.random.default_rng(seed), size or c.default_size)
            prog = c.build(tf)
            if not captured:  # compiled on first call with the actual extents: codegen mode compiles, then refuses to execute
                try:
                    prog(*inputs)
                except RuntimeError:
                    pass
            program = captured[-1]
            kernels = tf.get_all_generated_kernels()[seen:]
            seen += len(kernels)
            n_out = len(re.findall(r"\bout\[\d+\]\s*=", program.compiled_code()))
            print(f"[run_sim] {name}: {len(kernels)} kernels, {n_out} outputs", file=sys.stderr, flush=True)
            outs = run(build(program.compiled_code(), kernels, name), inputs, n_out, name)
            result[f"{spec}/coarsened"] = coarsened(kernels)
            for k, o in enumerate(outs):
                result[f"{spec}/{k}"] = o
    finally:
        os.dup2(saved, 1)
    np.savez(out_path, **result)
    print(f"[run_sim] {len(specs)} programs executed on the host from their emitted CUDA text")


if __name__ == "__main__":
    main()
