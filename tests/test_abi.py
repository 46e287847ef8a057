"""CPU tests of the drop-in boundary: libtfcuda.so loads without a GPU and exports exactly what include/tfcuda.h declares,
the ctypes mirror of the ABI structs has the reference's layout, and without a device every entry point fails loudly
(no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorfrost_b200 import abi  # noqa: E402

HEADER = os.path.join(ROOT, "include", "tfcuda.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    # function declarations only (typedef'd callback types have no tfcuda_ prefix)
    return sorted(set(re.findall(r"\b(tfcuda_[a-z0-9_]+)\s*\(", text)))


def test_header_and_ctypes_binding_agree():
    assert declared_symbols() == sorted(abi.EXPORTS), "include/tfcuda.h and tensorfrost_b200/abi.py list different entry points"


def test_library_exports_every_declared_symbol():
    lib = abi.lib()
    for name in declared_symbols():
        assert hasattr(lib, name), f"libtfcuda.so does not export {name}"


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "tfcuda.h"\nint main(void){ TFRuntime r; (void)r; return sizeof(TFDispatchInfo) == 64 ? 0 : 1; }\n')
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_struct_layouts_match_reference_abi():
    # Backend/TensorMemory.h:19-72 on LP64: sizes the generated host programs are compiled against
    assert C.sizeof(abi.TFDataFormat) == 16
    assert C.sizeof(abi.TFBuffer) == 40
    assert C.sizeof(abi.TFTensor) == 40
    assert C.sizeof(abi.TFDispatchInfo) == 64
    assert C.sizeof(abi.TFRuntime) == 56
    assert abi.TFTensor.format.offset == 8 and abi.TFTensor.dim.offset == 24 and abi.TFTensor.shape.offset == 32


def test_prelude_is_embedded_and_names_every_helper():
    prelude = abi.lib().tfcuda_prelude().decode()
    for helper in ("tf_min", "tf_clamp", "tf_lerp", "tf_smoothstep", "tf_sign", "tf_reversebits", "tf_pcg", "tf_pcgf", "tf_group_barrier",
                   "tf_atomic_add", "tf_atomic_add_prev", "tf_atomic_min", "tf_atomic_max", "tf_atomic_and", "tf_atomic_or", "tf_atomic_xor",
                   "asfloat", "asuint", "asint", "asbool"):
        assert helper in prelude, helper


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="a GPU is present: the no-device behaviour cannot be observed")
def test_no_device_means_loud_failure_not_fallback():
    lib = abi.lib()
    assert lib.tfcuda_init(-1) != 0
    assert b"no CPU fallback" in lib.tfcuda_last_error()
    assert lib.tfcuda_is_initialized() == 0
    # compute entry points refuse to run
    assert lib.tfcuda_reduce(0, 0, 1, 1, 1, 0, 0) != 0
    assert lib.tfcuda_radix_sort(1, 1, 0, 0, 4, 1, 32, 1) != 0
    assert lib.tfcuda_matmul(1, 1, 1, 1, 1, 1, 1, 2) != 0
    assert lib.tfcuda_malloc(16) == 0


def test_emitted_kernel_text_passes_nvrtc_without_a_device():
    """NVRTC cross-compiles for sm_100a on a CPU-only box: validates prelude + a kernel in the emitter's shape."""
    src = r'''
struct kernel_0_args { uint* mem[2]; uint var[2]; };
extern "C" __global__ void __launch_bounds__(256) kernel_0(const __grid_constant__ kernel_0_args tf_a)
{
  uint* out_mem = tf_a.mem[0];
  uint* in_mem = tf_a.mem[1];
  int var_n = asint(tf_a.var[0]);
  uint var__kernel_block_offset = asuint(tf_a.var[1]);
  int block_id = (int)(blockIdx.x + var__kernel_block_offset);
  int index_0 = block_id * 256 + (int)threadIdx.x;
  if (index_0 < var_n) {
    float x = asfloat(in_mem[tf_clamp(index_0, 0, var_n - 1)]);
    out_mem[index_0] = asuint(tf_lerp(tf_sin(x), tf_pcgf((uint)index_0), 0.5f));
    tf_atomic_add((uint*)out_mem, 0, (float)(x));
  }
}
'''
    rc = abi.lib().tfcuda_nvrtc_check(src.encode(), b"")
    assert rc == 0, abi.lib().tfcuda_last_error().decode()
