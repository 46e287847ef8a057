"""tensorfrost_b200 — a B200-native (sm_100a) CUDA execution backend for TensorFrost programs.

The product is two native artefacts:

* ``tensorfrost_b200/lib/libtfcuda.so`` — the C-ABI runtime declared in ``include/tfcuda.h``: device buffers, the
  ``TFRuntime`` callback table a compiled TensorFrost host program is called with, NVRTC compilation and dispatch of
  emitted kernels, the hand-written library kernels, the NCCL gradient exchange.  ``tensorfrost_b200.abi`` binds it
  with ctypes.
* ``build/tf_cuda/TensorFrost`` — the TensorFrost python module with the CUDA backend compiled in (reference frontend /
  IR / compiler unchanged + ``tensorfrost_b200/overlay``).  ``tensorfrost_b200.load()`` imports and initialises it;
  user code is unchanged apart from ``tf.initialize(tf.cuda)``.

There is no CPU fallback: without the native artefacts or without a CUDA device every entry point raises.
"""
import os
import sys

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(REPO_ROOT, "tensorfrost_b200", "lib", "libtfcuda.so")
MODULE_DIR = os.path.join(REPO_ROOT, "build", "tf_cuda")

_tf = None


def module_path():
    """Directory to put on sys.path so that `import TensorFrost` finds the CUDA-enabled module."""
    so = [f for f in os.listdir(os.path.join(MODULE_DIR, "TensorFrost"))] if os.path.isdir(os.path.join(MODULE_DIR, "TensorFrost")) else []
    if not any(f.startswith("TensorFrost") and f.endswith(".so") for f in so):
        raise ImportError(
            f"CUDA-enabled TensorFrost module not built under {MODULE_DIR}; run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(needs the reference sources) — there is no fallback module")
    return MODULE_DIR


def import_module():
    """Import the CUDA-enabled TensorFrost module without initialising a backend."""
    global _tf
    if _tf is None:
        path = module_path()
        if "TensorFrost" in sys.modules and not getattr(sys.modules["TensorFrost"], "__file__", "").startswith(path):
            raise ImportError("another TensorFrost module is already imported in this process (one backend module per process)")
        if path not in sys.path:
            sys.path.insert(0, path)
        import TensorFrost as tf  # noqa: E402
        if not hasattr(tf, "cuda"):
            raise ImportError(f"{tf.__file__} has no CUDA backend compiled in")
        _tf = tf
    return _tf


def load(kernel_compile_options=None):
    """`import TensorFrost as tf; tf.initialize(tf.cuda, options)` with the right module; returns tf.
    `kernel_compile_options`: extra NVRTC flags for emitted kernels (the reference's kernel_compile_options string), e.g.
    "-DTF_WARP_AGG_ATOMICS=0" or "--prec-div=false --prec-sqrt=false"; default: $TFCUDA_KERNEL_OPTIONS.

    Raises RuntimeError when no CUDA device is present (the backend never falls back to the CPU)."""
    if kernel_compile_options is None:
        kernel_compile_options = os.environ.get("TFCUDA_KERNEL_OPTIONS", "")
    tf = import_module()
    if str(tf.current_backend()) != str(tf.cuda):
        tf.initialize(tf.cuda, kernel_compile_options)
    return tf
