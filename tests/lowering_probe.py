"""Trace a few programs with the CUDA-enabled module in CODEGEN mode (no device) and print, as JSON, which library calls the
lowering produced.  Own process: the backend is a process-global singleton.  usage: python tests/lowering_probe.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def main():
    import tensorfrost_b200
    from tensorfrost_b200 import workloads
    tf = tensorfrost_b200.import_module()
    tf.initialize(tf.codegen, "", tf.cuda_lang)
    import cases
    out = {}
    keep = []
    seen = 0

    def collect(name, prog):
        nonlocal seen
        keep.append(prog)
        kernels = tf.get_all_generated_kernels()
        new = kernels[seen:]
        seen = len(kernels)
        text = [k[0][2] for k in new]
        markers = [t.split("tfcuda_lib:")[1].split(" ")[0] for t in text if "tfcuda_lib:" in t]
        out[name] = {"kernels": len(new), "library_calls": markers, "main": prog.get_main_function()}

    saved = os.dup(1)
    os.dup2(os.open(os.devnull, os.O_WRONLY), 1)
    try:
        collect("autograd_batched_dense", cases.CASES["autograd_batched_dense"].build(tf))
        collect("sort_pairs", workloads.compile_sort(tf, with_values=True))
        collect("matmul", workloads.compile_matmul(tf))
        collect("row_reductions", workloads.compile_row_reductions(tf, 8192))
        os.environ["TFCUDA_LIBRARY"] = "0"
        collect("row_reductions_generic", workloads.compile_row_reductions(tf, 8192))
        os.environ.pop("TFCUDA_LIBRARY")
    finally:
        os.dup2(saved, 1)
    print("PROBE " + json.dumps(out))


if __name__ == "__main__":
    main()
