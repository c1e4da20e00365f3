"""The C++ host mirror (include/hopfieldnetwork.hpp) compiles against the C ABI and behaves like the Go API:
CPU leg = builder validation + loud failure without a device; GPU leg (-m gpu) = a known-answer run."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_exe(tmp):
    from hopfield_network_go_b200 import build
    exe = os.path.join(tmp, "host_mirror_test")
    libdir = os.path.dirname(build.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"), "-o", exe,
                           "-L", libdir, "-lhopfield_b200", f"-Wl,-rpath,{libdir}"])
    return exe


def test_cpp_mirror_cpu(built, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the cpu leg asserts the no-device failure")
    out = subprocess.run([_build_exe(str(tmp_path)), "cpu"], capture_output=True, text=True)
    assert out.returncode == 0 and "OK cpu" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_mirror_gpu(built, tmp_path):
    out = subprocess.run([_build_exe(str(tmp_path)), "gpu"], capture_output=True, text=True)
    assert out.returncode == 0 and "OK gpu" in out.stdout, out.stdout + out.stderr
