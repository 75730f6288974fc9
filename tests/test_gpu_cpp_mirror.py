"""-m gpu: the reference's own unit tests / KATs for the path, restated in C++ against the host mirror
include/zkstark_b200.hpp (merkle.rs:112-182, prover.rs:42,64-78, main.rs:15-36), compiled with g++ and run."""
import os
import subprocess

import pytest

import zkstark_b200 as zk

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_reference_tests(tmp_path):
    exe = str(tmp_path / "reference_tests")
    libdir = os.path.dirname(zk.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "reference_tests.cpp"), "-o", exe,
                           "-L", libdir, "-lzkstark_b200", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "reference tests ok" in out.stdout


def test_cpp_prover_replay_per_function(tmp_path, transcripts):
    """prover.rs:31-292 call by call through the per-function exports with a host channel: golden transcripts at the
    reference's own configuration (sha256 b956f693...) and at 2^16, equality with zk_prove_fibsq, wire format."""
    exe = str(tmp_path / "prover_replay")
    libdir = os.path.dirname(zk.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "prover_replay.cpp"), "-o", exe,
                           "-L", libdir, "-lzkstark_b200", f"-Wl,-rpath,{libdir}"])
    assert transcripts["10"]["sha256"].startswith("b956f693")
    args = []
    for k in ("4", "10", "12", "16"):
        args += [k, transcripts[k]["sha256"]]
    out = subprocess.run([exe] + args, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "prover replay ok" in out.stdout
