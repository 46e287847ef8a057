"""Trace every case of tests/cases.py (and the fluid benchmark program) with the CUDA-enabled module in CODEGEN mode
(no device needed: tf.initialize(tf.codegen, "", tf.cuda_lang)), collect the CUDA C++ the emitter produced for every fused
kernel and compile it with NVRTC for sm_100a through the C-ABI (tfcuda_nvrtc_check).  Run in its own process: the
backend is a process-global singleton.

usage: python tests/emit_all.py [case ...]        exit code 0 when every kernel compiles
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def main():
    import tensorfrost_b200
    from tensorfrost_b200 import abi, workloads
    tf = tensorfrost_b200.import_module()
    tf.initialize(tf.codegen, "", tf.cuda_lang)
    import cases
    names = sys.argv[1:] or (sorted(cases.CASES) + ["fluid"])
    lib = abi.lib()
    failures = []
    checked = 0
    seen = 0
    keep = []  # programs must stay alive: the kernel registry holds raw pointers into them (Backend/KernelManager.cpp:4-8)
    for name in names:
        if name == "fluid":
            keep.append(workloads.load_fluid(tf, 256, 256))
        else:
            c = cases.CASES[name]
            prog = c.build(tf)
            keep.append(prog)
            if not hasattr(prog, "get_kernels"):
                # programs specialised on an input extent are built on first call; codegen mode cannot execute, so build directly
                try:
                    prog(*c.make_inputs(np.random.default_rng(0), c.default_size))
                except RuntimeError:
                    pass
        kernels = tf.get_all_generated_kernels()
        new = kernels[seen:]
        seen = len(kernels)
        # one NVRTC unit per case, like the runtime's chunks
        src = "\n".join(k[0][1] + k[0][2] for k in new)
        rc = lib.tfcuda_nvrtc_check(src.encode(), b"")
        checked += len(new)
        if rc != 0:
            failures.append((name, lib.tfcuda_last_error().decode(errors="replace")[:3000]))
            print(f"[emit] {name}: {len(new)} kernels FAILED", flush=True)
        else:
            print(f"[emit] {name}: {len(new)} kernels ok", flush=True)
    for name, log in failures:
        print(f"---- {name} ----\n{log}")
    print(f"[emit] {checked} kernels checked, {len(failures)} cases failed")
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
