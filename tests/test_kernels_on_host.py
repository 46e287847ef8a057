"""CPU test: hand-written CUDA kernels compiled for the host through tests/cpu_sim/cuda_host_shim.h and run with one host thread per
CUDA thread of a block around a std::barrier (tests/cpu_sim/kernel_on_host.cpp), against a float64 reference.

Subject: the skinny matmul of csrc/matmul_rows.cu, which was written after round 1's GPU budget ended and has never run on hardware -
every template configuration its dispatcher uses, NCA's four shapes, ragged tails, scalar fallbacks.  Control: the weight-gradient
kernel of csrc/matmul_tn.cu, which IS validated on hardware, through the same harness.  Also the n-body step: the scalar kernel
(hardware-validated) and the packed f32x2 kernel at the sizes of tests/test_zy_late_gpu.py, with host stand-ins for its PTX wrappers."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_hand_written_kernels_run_correctly_on_the_host(tmp_path):
    exe = str(tmp_path / "kernel_on_host")
    cmd = ["g++", "-std=c++20", "-O2", "-pthread", "-w", "-I", os.path.join(HERE, "cpu_sim"), "-I", os.path.join(ROOT, "tensorfrost_b200", "csrc"),
           os.path.join(HERE, "cpu_sim", "kernel_on_host.cpp"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-3000:] + r.stderr[-1000:]
    assert r.stdout.count(" ok") == 21 and "FAIL" not in r.stdout
