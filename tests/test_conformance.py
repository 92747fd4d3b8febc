"""include/sped.h is valid C and the library can be driven from C in the reference's call order."""
import os
import subprocess

import pytest

from spin_ed_b200 import ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "conformance")
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "conformance.c"), "-o", exe, "-L", os.path.dirname(ffi.LIB_PATH),
                           "-lsped", "-lm", "-Wl,-rpath," + os.path.dirname(ffi.LIB_PATH)])
    return exe


def test_header_compiles_as_c_and_links(tmp_path):
    _build(tmp_path)


@pytest.mark.gpu
def test_c_driver_runs_the_reference_call_sequence(tmp_path):
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert out.stdout.startswith("CONFORMANCE_OK 13 -18.06178541")
