"""The drop-in boundary without Python in the loop: examples/c_api_demo.c is compiled with gcc against
include/ndsb200.h and libndsb200.so and run as a stand-alone process."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp_path):
    exe = os.path.join(str(tmp_path), "c_api_demo")
    lib = os.path.join(ROOT, "ndsimulator_b200", "csrc")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_api_demo.c"), "-o", exe, "-L", lib, "-lndsb200",
                           f"-Wl,-rpath,{lib}", "-lm"])
    return exe


def test_c_demo_compiles_against_the_header(tmp_path):
    build(tmp_path)   # plain C99 + the public header only: no torch / CUDA types leak through the boundary


@pytest.mark.gpu
def test_c_demo_runs_c2_physics(tmp_path):
    out = subprocess.run([build(tmp_path), "65536", "4000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "md_block<f32x2,nd2,Mueller2d,Original,langevin,nobias>" in out.stdout
    T = float(re.search(r"<T> = ([0-9.]+) K", out.stdout).group(1))
    assert abs(T - 599.6) < 12.0           # T_eff = T (1 + c1)^2 / (1 + c1^2)
    assert "steps 4000 launches" in out.stdout
