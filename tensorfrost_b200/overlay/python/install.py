#!/usr/bin/env python3
"""Installs the backend-aware python helpers into the built TensorFrost package (build/tf_cuda/TensorFrost).

The reference's `tf.sort.radix` (Python/TensorFrost/sort.py:38-187) already branches on `tf.current_backend()` (sort.py:122);
this adds the CUDA branch: on `tf.cuda`, a 1-D sort becomes ONE library call in the traced program (tf.cuda_library_sort ->
tfcuda_radix_sort) instead of the 13-kernel / 37-dispatch generic pipeline.  Same contract: stable ascending LSD radix over the
low `max_bits` bits of the mapped key (float / int key bijections), returning keys or (keys, values).  `bits_per_pass` is a tuning
knob of the generic algorithm and is ignored (the library uses 8-bit digits).  TFCUDA_LIBRARY=0 keeps the generic path.
The edit is an append to the package's own sort.py in the BUILD directory; nothing of the reference is stored in this repo.

usage: install.py <package dir>
"""
import sys
from pathlib import Path

APPEND = '''

# ---- tensorfrost_b200: CUDA backend dispatch (appended by overlay/python/install.py) --------------------------------
import os as _os

_radix_generic = radix


def radix(keys, values=None, bits_per_pass=6, max_bits=32):
    if tf.cuda_library_active() and len(keys.shape) == 1 and _os.environ.get("TFCUDA_LIBRARY", "1") != "0":
        if values is not None and (len(values.shape) != 1 or values.type == tf.bool1):
            return _radix_generic(keys, values, bits_per_pass, max_bits)
        tf.region_begin('Radix sort')
        out = tf.cuda_library_sort(keys, values, max_bits)
        tf.region_end('Radix sort')
        return out
    return _radix_generic(keys, values, bits_per_pass, max_bits)
'''


def main():
    pkg = Path(sys.argv[1])
    sort_py = pkg / "sort.py"
    text = sort_py.read_text()
    if "tensorfrost_b200: CUDA backend dispatch" not in text:
        sort_py.write_text(text + APPEND)
    print("install.py: patched", sort_py)


if __name__ == "__main__":
    main()
