"""TEST INFRASTRUCTURE.  The drop-in boundary as an executable: a TensorFrost program is traced in CODEGEN mode
(no backend executes anything), the host program text the reference generated is compiled with g++ exactly as
Backends/CPU/KernelCompiler.cpp would, the emitted CUDA kernels are handed to `tfcuda_compile_kernels`, and the program's
`main(in, out, TFRuntime)` is called with the callback table `tfcuda_runtime()` exports - alloc / dealloc / readback / writeback /
dispatch / region all run inside libtfcuda.so on the B200, through ctypes only.  Outputs go to an .npz for the caller to compare with
the reference's golden fixtures.

usage: python tests/standalone/run_standalone.py <out.npz> <case[:size[:seed]]> [...]       (own process)
       --dry: stop before anything needs a device (trace, emit, g++ the host program): what the CPU suite runs
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

TYPE_ID = {"f": 0, "u": 1, "i": 2, "b": 3}  # TFType (Backend/TensorMemory.h / Operations.h)
NP_OF = {0: np.float32, 1: np.uint32, 2: np.int32, 3: np.uint32}


def kernel_sources(kernels, host_code, abi):
    """TFCudaKernelSource records for every emitted kernel the host program dispatches (group sizes are baked into the kernel; the host
    text carries the IR's as the last argument of tf.dispatch, CPP.cpp:622-629, the kernel text the launched one as `// tfcuda_block:`)."""
    groups = {}
    for line in host_code.splitlines():
        m = re.search(r"tf\.dispatch\((\d+),.*\{([^{}]*)\}\);\s*$", line)
        if m:
            g = [int(x) for x in m.group(2).replace(" ", "").split(",") if x]
            groups[int(m.group(1))] = (g + [1, 1, 1])[:3]
    records, keep = [], []
    for k in kernels:
        src = k[0][1] + k[0][2]
        m = re.search(r"void (?:__launch_bounds__\(\d+\) )?kernel_(\d+)\(", src)
        if not m:
            raise RuntimeError("a kernel without emitted text (library call?) - trace with TFCUDA_LIBRARY=0")
        kid = int(m.group(1))
        if kid not in groups:
            continue
        nm = re.search(r"uint\* mem\[(\d+)\];", src)
        nv = re.search(r"uint var\[(\d+)\];", src)
        rec = abi.TFCudaKernelSource()
        rec.kernel_id = kid
        entry, text = f"kernel_{kid}".encode(), src.encode()
        keep += [entry, text]
        rec.entry, rec.source = entry, text
        # threads per block: the emitter states them (a coarsened kernel is launched with a smaller block than the one the host
        # program's block count was computed for); the tf.dispatch text is the IR's block
        lb = re.search(r"// tfcuda_block: (\d+) (\d+) (\d+)", src)
        block = [int(x) for x in lb.groups()] if lb else groups[kid]
        for d in range(3):
            rec.group[d] = block[d]
        rec.n_mem = int(nm.group(1)) if nm else 0
        rec.n_var = int(nv.group(1))
        rec.library_op = 0
        records.append(rec)
    return records, keep


def build_host(host_code, tag):
    work = tempfile.mkdtemp(prefix=f"tfstandalone_{tag}_")
    src = os.path.join(work, "host.cpp")
    with open(src, "w") as f:
        f.write("#define main tf_program_main\n" + host_code + "\n" + open(os.path.join(HERE, "standalone_entry.inc")).read())
    so = os.path.join(work, "host.so")
    r = subprocess.run(["g++", "-O1", "-w", "-shared", "-fPIC", "-std=c++17", src, "-o", so], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed for {tag}:\n{r.stderr[-3000:]}")
    return so


def to_words(a):
    a = np.asarray(a)
    if a.dtype == np.bool_:
        return np.ascontiguousarray(a.astype(np.uint32)), 3
    if a.dtype.kind == "f":
        return np.ascontiguousarray(a.astype(np.float32)).view(np.uint32), 0
    if a.dtype.kind == "u":
        return np.ascontiguousarray(a.astype(np.uint32)), 1
    return np.ascontiguousarray(a.astype(np.int32)).view(np.uint32), 2


def run_on_device(abi, so, records, inputs, n_out, tag):
    lib = abi.lib()
    if records:
        arr = (abi.TFCudaKernelSource * len(records))(*records)
        abi.check(lib.tfcuda_compile_kernels(arr, len(records), b""), f"tfcuda_compile_kernels({tag})")
    host = C.CDLL(so)
    host.standalone_run.argtypes = [C.POINTER(abi.TFTensor), C.POINTER(abi.TFTensor), abi.TFRuntime, C.c_char_p, C.c_size_t]
    host.standalone_run.restype = C.c_int
    ins = (abi.TFTensor * max(len(inputs), 1))()
    keep, buffers = [], []
    for i, a in enumerate(inputs):
        words, type_id = to_words(a)
        buf = lib.tfcuda_buffer_create(max(words.size, 1))
        if not buf:
            raise RuntimeError(lib.tfcuda_last_error().decode())
        buffers.append(buf)
        abi.check(lib.tfcuda_buffer_write(buf, 0, words.ctypes.data, words.size), "tfcuda_buffer_write")
        shape = (C.c_size_t * max(words.ndim, 1))(*words.shape)
        keep.append(shape)
        ins[i].buffer = buf
        ins[i].format.type, ins[i].format.size = type_id, 32
        ins[i].dim = words.ndim
        ins[i].shape = C.cast(shape, C.POINTER(C.c_size_t))
    outs = (abi.TFTensor * max(n_out, 1))()
    err = C.create_string_buffer(4096)
    rc = host.standalone_run(ins, outs, lib.tfcuda_runtime(), err, len(err))
    if rc != 0:
        raise RuntimeError(f"{tag}: the program failed through the TFRuntime table: {err.value.decode(errors='replace')}")
    abi.check(lib.tfcuda_sync(), "tfcuda_sync")
    result = []
    for i in range(n_out):
        t = outs[i]
        shape = tuple(t.shape[d] for d in range(t.dim))
        n = int(np.prod(shape)) if shape else 1
        words = np.empty(n, np.uint32)
        abi.check(lib.tfcuda_buffer_read(t.buffer, 0, words.ctypes.data, n), "tfcuda_buffer_read")
        a = words.reshape(shape).view(NP_OF[t.format.type])
        result.append(a != 0 if t.format.type == 3 else a)
    for b in buffers:
        lib.tfcuda_buffer_destroy(b)
    return result


def main():
    argv = [a for a in sys.argv[1:] if a != "--dry"]
    dry = "--dry" in sys.argv
    out_path, specs = argv[0], argv[1:]
    os.environ["TFCUDA_LIBRARY"] = "0"  # library calls are dispatched by the in-module glue, not by libtfcuda's own table
    import tensorfrost_b200
    from tensorfrost_b200 import abi
    tf = tensorfrost_b200.import_module()
    saved = os.dup(1)
    os.dup2(os.open(os.devnull, os.O_WRONLY), 1)
    result = {}
    try:
        tf.initialize(tf.codegen, "", tf.cuda_lang)
        if not dry:
            abi.init(-1)
        import cases
        captured, keep_programs = [], []
        real_compile = tf.compile

        def capturing_compile(fn):
            p = real_compile(fn)
            captured.append(p)
            keep_programs.append(p)
            return p
        tf.compile = capturing_compile
        seen = 0
        for spec in specs:
            parts = spec.split(":")
            name = parts[0]
            size = int(parts[1]) if len(parts) > 1 and parts[1] else None
            seed = int(parts[2]) if len(parts) > 2 else 0
            c = cases.CASES[name]
            inputs = c.make_inputs(np.random.default_rng(seed), size or c.default_size)
            del captured[:]
            prog = c.build(tf)
            if not captured:
                try:
                    prog(*inputs)
                except RuntimeError:
                    pass
            program = captured[-1]
            kernels = tf.get_all_generated_kernels()[seen:]
            seen += len(kernels)
            host_code = program.compiled_code()
            n_out = len(re.findall(r"\bout\[\d+\]\s*=", host_code))
            records, keep = kernel_sources(kernels, host_code, abi)
            so = build_host(host_code, name)
            print(f"[standalone] {name}: {len(records)} kernels, {n_out} outputs, host program {so}", file=sys.stderr, flush=True)
            if dry:
                result[f"{spec}/kernels"] = np.array(len(records))
                continue
            outs = run_on_device(abi, so, records, inputs, n_out, name)
            for k, o in enumerate(outs):
                result[f"{spec}/{k}"] = o
    finally:
        os.dup2(saved, 1)
    np.savez(out_path, **result)
    print(f"[standalone] {len(specs)} programs {'prepared' if dry else 'executed through tfcuda_runtime()'}")


if __name__ == "__main__":
    main()
