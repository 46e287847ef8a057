"""Why do the uploads and downloads of bench.py's pipelined end-to-end loop not overlap each other?  Variants of the loop, timed."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tensorfrost_b200  # noqa: E402
from tensorfrost_b200 import workloads  # noqa: E402

tf = tensorfrost_b200.load()
n = 2048
fluid = workloads.load_fluid(tf, n, n)
host_inputs = workloads.fluid_inputs(n, n)
pin_in = [tf.cuda_pinned_array([n, n], "float32") for _ in range(4)]
pin_out = [[tf.cuda_pinned_array([n, n], "float32") for _ in range(4)] for _ in range(2)]
sets = [[tf.cuda_tensor(a) for a in host_inputs] for _ in range(2)]


def scale():
    a = tf.input([-1, -1], tf.float32)
    return a * 1.0001


trivial = tf.compile(scale)


def loop(steps, compute, downloads=True, uploads=True):
    if uploads:
        for k in range(4):
            tf.cuda_upload_async(sets[0][k], pin_in[k])
    for s in range(steps):
        cur, nxt = sets[s % 2], sets[(s + 1) % 2]
        if uploads:
            tf.cuda_wait_uploads()
            if s + 1 < steps:
                for k in range(4):
                    tf.cuda_upload_async(nxt[k], pin_in[k])
        if compute == "fluid":
            out, _ = workloads.fluid_step(fluid, cur)
        elif compute == "trivial":
            out = [trivial(cur[k]) for k in range(4)]
        else:
            out = cur
        if downloads:
            for k in range(4):
                tf.cuda_download_async(out[k], pin_out[s % 2][k])
    tf.cuda_copy_sync()
    tf.cuda_synchronize()


def timed(name, **kw):
    loop(4, **kw)
    t0 = time.perf_counter()
    loop(20, **kw)
    ms = (time.perf_counter() - t0) / 20 * 1e3
    print(f"{name:44s} {ms:7.3f} ms/step", flush=True)


print("TFCUDA_COPY_HOSTFUNC =", os.environ.get("TFCUDA_COPY_HOSTFUNC", "1"), flush=True)
timed("uploads only", compute=None, downloads=False)
timed("downloads only", compute=None, uploads=False)
timed("uploads + downloads, no kernels", compute=None)
timed("uploads + trivial kernels + downloads", compute="trivial")
timed("uploads + fluid step + downloads (bench)", compute="fluid")
timed("fluid step + downloads", compute="fluid", uploads=False)
timed("uploads + fluid step", compute="fluid", downloads=False)
