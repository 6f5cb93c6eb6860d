"""Builds and runs tests/cpp/test_host_mirror.cpp: the reference's own unit tests restated against the C++ host
mirror (gnss-sdr-rs_b200/host/gnss_sdr_rs.hpp) over the C-ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, ffi):
    exe = str(tmp_path / "test_host_mirror")
    libdir = os.path.dirname(ffi.LIB_PATH)
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp"), "-o", exe,
                    "-L", libdir, "-lgnss_b200", "-lpthread", "-Wl,-rpath," + libdir], check=True)
    return exe


def test_cpp_host_mirror_cpu_part(tmp_path, ffi):
    r = subprocess.run([_build(tmp_path, ffi), "--cpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_static_archive_links(tmp_path, ffi):
    """the reference's build.rs links a static archive from src/c_lib (libconvenience.a); libgnss_b200.a is the
    same objects as the shared library and must link with cudart_static alone (no NCCL: it is dlopen'ed on demand)."""
    libdir = os.path.dirname(ffi.LIB_PATH)
    arch = os.path.join(libdir, "libgnss_b200.a")
    assert os.path.exists(arch), "build.sh did not produce the static archive"
    cuda_lib = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "lib64")
    exe = str(tmp_path / "test_host_mirror_static")
    subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp"), "-o", exe, arch,
                    "-L", cuda_lib, "-lcudart_static", "-ldl", "-lrt", "-lpthread"], check=True)
    r = subprocess.run([exe, "--cpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_host_mirror_reference_tests(tmp_path, ffi):
    r = subprocess.run([_build(tmp_path, ffi)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok" in r.stdout


def test_fmod_small_is_bit_exact(tmp_path):
    """The tracking kernel's exact fmod (one FMA from |x|, quotient fixed up by the remainder's sign/size) against glibc
    fmodf on 2e7 arguments in the ranges the loops produce, including values adjacent to exact multiples."""
    exe = str(tmp_path / "test_fmod_small")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-frounding-math", os.path.join(ROOT, "tests", "cpp", "test_fmod_small.c"),
                    "-o", exe, "-lm"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "bad=0" in r.stdout, r.stdout
