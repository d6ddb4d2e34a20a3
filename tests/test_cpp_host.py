"""The C++ host-side mirror (include/mtt_b200.hpp) compiles against the C-ABI with plain g++ and
reproduces the reference's known-answer test through it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_test_kat")


def _build():
    import multitypetree_b200 as mtt
    mtt.lib()
    libdir = os.path.join(ROOT, "multitypetree_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_kat.cpp"), "-o", EXE, "-L", libdir,
           "-l:libmtt_b200.so", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_cpp_mirror_compiles_and_fails_loudly_without_gpu():
    exe = _build()
    r = subprocess.run([exe, "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "no device" in r.stdout or "a device is present" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_reference_kat():
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    kat = float([l for l in r.stdout.splitlines() if l.startswith("KAT")][0].split()[1])
    assert abs(kat - (-16.52831)) < 1e-5
    assert "INVALID -inf" in r.stdout
