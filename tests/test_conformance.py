"""include/sped.h is valid C and the library can be driven from C in the reference's call order."""
import os
import subprocess

import pytest

from spin_ed_b200 import ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, source="conformance.c"):
    exe = str(tmp_path / source[:-2])
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", source), "-o", exe, "-L", os.path.dirname(ffi.LIB_PATH),
                           "-lsped", "-lm", "-Wl,-rpath," + os.path.dirname(ffi.LIB_PATH)])
    return exe


def test_header_compiles_as_c_and_links(tmp_path):
    _build(tmp_path)
    _build(tmp_path, "conformance_full.c")  # takes the address of all 32 ls_* symbols: a missing export is a link error


@pytest.mark.parametrize("destroy", [False, True])
def test_process_may_exit_while_nvrtc_compiles(tmp_path, destroy):
    """ls_build starts NVRTC on a background thread; a process that ends at once -- basis destroyed or
    leaked -- must exit cleanly (seen once as SIGILL in the C conformance driver before the
    compilation was waited for)."""
    exe = _build(tmp_path, "exit_race.c")
    env = dict(os.environ, SPED_JIT_PREFETCH="1", SPED_CACHE_DIR="")
    for _ in range(3):
        out = subprocess.run([exe] + (["destroy"] if destroy else []), capture_output=True, text=True, timeout=300, env=env)
        assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
        assert "build rc" in out.stdout


@pytest.mark.gpu
def test_c_driver_runs_the_reference_call_sequence(tmp_path):
    out = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert out.stdout.startswith("CONFORMANCE_OK 13 -18.06178541")


@pytest.mark.gpu
def test_c_driver_covers_all_symbols_and_every_destroy_order(tmp_path):
    """all 32 ls_* imports of Internal.hs, ls_build_unsafe, 1-/3-/4-site and complex terms, and the six
    destroy orders of {basis, operator, states} (GHC finalizers run in no particular order)"""
    out = subprocess.run([_build(tmp_path, "conformance_full.c")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert out.stdout.startswith("CONFORMANCE_FULL_OK 10")
