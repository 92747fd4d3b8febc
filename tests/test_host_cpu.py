"""CPU-side tests (-m "not gpu"): the library loads and exports the ABI, the host logic of the
product (group closure, error behaviour, Burnside dimension, canonicalisation-program compiler,
projected eigen-solver, row partition) agrees with the oracle / numpy."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import SMALL_DECKS, extra_configs, oracle_problem, product_problem
from spin_ed_b200 import config, decks, ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    declared = set()
    for hdr in ("sped.h", "sped_selftest.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        declared |= set(re.findall(r"\b((?:ls|sped)_[a-z0-9_]+)\s*\(", text))
    declared -= {"sped_monitor_fn"}
    assert len(declared) >= 50
    L = C.CDLL(ffi.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(ffi.SIGNATURES), declared ^ set(ffi.SIGNATURES)
    # the host emulation of the kernels is a test-only library, not part of the product
    assert not hasattr(L, "sped_selftest_emulate_matvec")
    assert hasattr(C.CDLL(ffi.EMUL_LIB_PATH), "sped_selftest_emulate_matvec")


def test_symmetry_semantics_of_reference_spec():
    # /root/reference/test/Spec.hs:38-43
    s = ffi.mkSymmetry([3, 2, 1, 0], 0)
    assert ffi.getSector(s) == 0 and ffi.getPeriodicity(s) == 2 and ffi.getPhase(s) == 0.0
    with pytest.raises(ffi.LatticeSymmetriesException):
        ffi.mkSymmetry([4, 3, 4, 1], 0)
    with pytest.raises(ffi.LatticeSymmetriesException):
        ffi.mkSymmetry([4, 3, 2, 1, 0], 3)
    with pytest.raises(ffi.SpinEDException):
        ffi.mkSymmetry([1, -1, 0], 0)
    assert ffi.getPhase(ffi.mkSymmetry([1, 2, 3, 0], 1)) == 0.25


def test_group_semantics_of_reference_spec():
    # /root/reference/test/Spec.hs:72-79 (the spin flip generator of that stale test is now the
    # basis' spin_inversion; the lattice part D4 has 8 elements)
    s1 = ffi.mkSymmetry([3, 2, 1, 0], 0)
    s2 = ffi.mkSymmetry([1, 2, 3, 0], 0)
    s2b = ffi.mkSymmetry([1, 2, 3, 0], 1)
    assert ffi.getGroupSize(ffi.mkGroup([s1, s2])) == 8
    with pytest.raises(ffi.LatticeSymmetriesException) as e:
        ffi.mkGroup([s1, s2b])
    assert e.value.eCode == 11
    assert ffi.getGroupSize(ffi.mkGroup([])) == 1


def test_basis_argument_errors():
    g = ffi.mkGroup([])
    with pytest.raises(ffi.SpinEDException):
        ffi.mkBasis(g, -2)
    with pytest.raises(ffi.SpinEDException):
        config.toBasis(config.BasisSpec(100, None, None, []))
    with pytest.raises(ffi.LatticeSymmetriesException):
        ffi.mkBasis(g, 4, 5)
    with pytest.raises(ffi.LatticeSymmetriesException):
        ffi.mkBasis(g, 4, 1, 1)
    b = ffi.mkBasis(g, 4, 2)
    with pytest.raises(ffi.LatticeSymmetriesException) as e:
        ffi.getNumberStates(b)  # not built
    assert e.value.eCode == 14


def test_interaction_semantics():
    # /root/reference/test/Spec.hs:96-104
    t = ffi.mkInteraction([[1.0, 0.0], [0.0, -1.0]], [[0], [2]])
    assert ffi.isRealInteraction(t)
    t2 = ffi.mkInteraction([[0, [0, -1]], [[0, 1], 0]], [[0]])
    assert not ffi.isRealInteraction(t2)
    with pytest.raises(ffi.SpinEDException):
        ffi.mkInteraction([[1, 2, 1], [4, 2, 1], [1, -2, 3]], [[0, 0, 0]])
    with pytest.raises(ffi.SpinEDException):
        ffi.mkInteraction([[1, 0], [0, 1]], [])


def test_compute_fails_loudly_without_gpu():
    try:
        ffi.deviceCount()
        pytest.skip("a CUDA device is present")
    except ffi.LatticeSymmetriesException as e:
        assert e.eCode == 100
    uc = product_problem(decks.load("heisenberg_chain_10"))
    with pytest.raises(ffi.LatticeSymmetriesException) as e:
        ffi.buildBasis(uc.cBasis)
    assert e.value.eCode == 100


@pytest.mark.parametrize("name", SMALL_DECKS + ["heisenberg_square_5x5", "heisenberg_pyrochlore_32", "heisenberg_square_6x6",
                                                "heisenberg_chain_40", "heisenberg_chain_42", "chain_50_hw2", "chain_64_hw2",
                                                "chain_40_hw3_k", "chain_10_inv_nohw", "chain_8_k1_complex"])
def test_program_matches_oracle_state_info(oracle, name):
    """The compiled canonicalisation program (interpreted on the host) must give the oracle's
    representative, character and norm for random states of the sector."""
    cfg = extra_configs()[name] if name in extra_configs() else decks.load(name)
    ob, _ = oracle_problem(oracle, cfg)
    if ob.group_size * (2 if ob.spin_inversion else 1) <= 1:
        pytest.skip("trivial group")
    uc = product_problem(cfg)
    n, hw = ob.number_spins, ob.hamming_weight
    rng = np.random.default_rng(1234)
    states = []
    for _ in range(300):
        if hw is None:
            states.append(int(rng.integers(0, 1 << n, dtype=np.uint64)))
        else:
            pos = rng.choice(n, size=hw, replace=False)
            states.append(int(sum(1 << int(p) for p in pos)))
    s = np.array(states, dtype=np.uint64)
    reps = np.zeros(len(s), dtype=np.uint64)
    phases = np.zeros(len(s), dtype=np.int32)
    stabs = np.zeros(len(s), dtype=np.int32)
    ffi.checkStatus(ffi.lib().sped_selftest_program(uc.cBasis._ptr, len(s), s.ctypes.data, reps.ctypes.data,
                                                    phases.ctypes.data, stabs.ctypes.data))
    order = ob.group_size * (2 if ob.spin_inversion else 1)
    denom = np.lcm.reduce([2] + [oracle.periodicity(sym["permutation"]) for sym in cfg["basis"]["symmetries"]])
    for x, r, ph, st in zip(states, reps, phases, stabs):
        orep, ochi, onorm = ob.state_info(x)
        assert int(r) == orep
        assert abs(np.sqrt(st / order) - onorm) < 1e-15
        if onorm > 0:
            assert abs(np.exp(2j * np.pi * ph / denom) - ochi) < 1e-12


@pytest.mark.parametrize("name", sorted(decks.names()))
def test_burnside_dimension(oracle, name):
    cfg = decks.load(name)
    uc = product_problem(cfg)
    out = C.c_uint64(0)
    ffi.checkStatus(ffi.lib().sped_selftest_burnside(uc.cBasis._ptr, C.byref(out)))
    known = {"heisenberg_chain_4": 16, "heisenberg_chain_10": 13, "heisenberg_kagome_12": 924, "heisenberg_square_4x4": 107,
             "heisenberg_triangular_19": 4862, "xxz_triangular_19": 524288, "heisenberg_chain_24": 2704156,
             "heisenberg_square_6x6": 15804956, "heisenberg_pyrochlore_32": 789438, "heisenberg_chain_40": 861725794,
             "heisenberg_chain_42": 3204236779, "heisenberg_square_5x5": 208012}
    assert out.value == known[name]


@pytest.mark.parametrize("name", sorted(extra_configs()))
def test_burnside_matches_oracle_on_corner_cases(oracle, name):
    cfg = extra_configs()[name]
    ob, _ = oracle_problem(oracle, cfg)
    ob.build()
    uc = product_problem(cfg)
    out = C.c_uint64(0)
    ffi.checkStatus(ffi.lib().sped_selftest_burnside(uc.cBasis._ptr, C.byref(out)))
    assert out.value == ob.number_states


def test_program_is_cheap_for_translation_groups():
    uc = product_problem(decks.load("heisenberg_square_6x6"))
    st = ffi.basisProgramStats(uc.cBasis)
    assert st["steps"] == 288
    # translations are two masked rotates; a generic Benes network would be ~11 swaps per element
    assert st["rotate_mask_ops"] + st["delta_swap_ops"] < 4 * st["steps"]


@pytest.mark.parametrize("m", [1, 2, 5, 17, 40])
@pytest.mark.parametrize("cplx", [False, True])
def test_small_eigh_matches_numpy(m, cplx):
    rng = np.random.default_rng(m)
    A = rng.standard_normal((m, m)) + (1j * rng.standard_normal((m, m)) if cplx else 0)
    A = (A + A.conj().T) / 2
    a = np.ascontiguousarray(A.astype(np.complex128))
    ev = np.zeros(m)
    V = np.zeros((m, m), dtype=np.complex128)
    ffi.checkStatus(ffi.lib().sped_selftest_small_eigh(m, a.ctypes.data, ev.ctypes.data, V.ctypes.data))
    assert np.allclose(ev, np.linalg.eigvalsh(A), atol=1e-12)
    assert np.allclose(A @ V, V * ev, atol=1e-11)
    assert np.allclose(V.conj().T @ V, np.eye(m), atol=1e-12)


@pytest.mark.parametrize("kind", ["diagonal", "degenerate", "arrow", "graded", "identity"])
def test_small_eigh_real_fast_path_on_structured_matrices(kind):
    """The projected matrices of the solver are far from generic: diagonal Ritz blocks bordered by
    new rows (arrow shape), exactly degenerate levels, entries spread over many orders of magnitude."""
    rng = np.random.default_rng(3)
    m = 24
    if kind == "diagonal":
        A = np.diag(np.sort(rng.standard_normal(m)))
    elif kind == "identity":
        A = np.eye(m)
    elif kind == "degenerate":
        q, _ = np.linalg.qr(rng.standard_normal((m, m)))
        A = q @ np.diag(np.repeat([-8.0, -3.0, 0.0, 5.0], 6)) @ q.T
    elif kind == "arrow":
        A = np.diag(np.sort(rng.standard_normal(m)) * 10)
        A[:, -3:] = 1e-7 * rng.standard_normal((m, 3))
        A[-3:, :] = A[:, -3:].T
        A[-3:, -3:] = np.diag([1.0, 2.0, 3.0])
    else:
        s = 10.0 ** rng.uniform(-9, 2, m)
        B = rng.standard_normal((m, m))
        A = s[:, None] * (B + B.T) * s[None, :]
    A = (A + A.T) / 2
    a = np.ascontiguousarray(A.astype(np.complex128))
    ev = np.zeros(m)
    V = np.zeros((m, m), dtype=np.complex128)
    ffi.checkStatus(ffi.lib().sped_selftest_small_eigh(m, a.ctypes.data, ev.ctypes.data, V.ctypes.data))
    scale = max(1.0, np.abs(A).max())
    assert np.all(np.diff(ev) >= 0)
    assert np.allclose(ev, np.linalg.eigvalsh(A), atol=1e-12 * scale)
    assert np.allclose(A @ V, V * ev, atol=1e-11 * scale)
    assert np.allclose(V.conj().T @ V, np.eye(m), atol=1e-12)
    assert np.all(V.imag == 0)


def test_row_distribution_is_a_balanced_block_cyclic_partition():
    import ctypes as C

    for n in [0, 1, 7, 13, 1000, 70001, 15804956]:
        for world in [1, 2, 3, 8]:
            seen = np.zeros(n, dtype=np.int64)
            counts = []
            chunk = None
            for r in range(world):
                d = ffi.rowDistribution(n, world, r)
                chunk = d.chunk if chunk is None else chunk
                assert d.chunk == chunk and d.n == n and d.world == world and d.rank == r
                counts.append(d.n_local)
                if n <= 70001:
                    rows = d.local_rows()
                    assert len(rows) == d.n_local and np.all(np.diff(rows.astype(np.int64)) > 0)
                    seen[rows.astype(np.int64)] += 1
                    pos = d.global_to_position(rows)
                    assert np.array_equal(pos, np.uint64(r * d.chunk) + np.arange(d.n_local, dtype=np.uint64))
                    for i in (0, d.n_local // 2, d.n_local - 1):  # the C functions agree with the numpy mirror
                        if d.n_local:
                            g = ffi.lib().sped_dist_local_to_global(C.byref(d), int(i))
                            assert g == int(rows[i])
                            assert ffi.lib().sped_dist_global_to_position(C.byref(d), g) == r * d.chunk + i
            assert sum(counts) == n and max(counts) == (chunk if world > 1 or n else 0) or n == 0 or world == 1
            if n <= 70001:
                assert np.all(seen == 1)
            if n >= 1000:  # balanced to within one block
                assert max(counts) - min(counts) <= (1 << ffi.rowDistribution(n, world, 0).log2_block)


def test_row_distribution_at_the_sharded_deck_sizes():
    """chain_40 / chain_42 on 2..8 ranks: large blocks (near-diagonal elements stay on the owning
    rank), at least 256 blocks per rank, balance within 0.1 %, and the replicated vector still
    addressable with 32-bit positions (the operator cache stores u32 positions)."""
    import ctypes as C

    for n in (861725794, 3204236779):
        for world in (2, 4, 8):
            dists = [ffi.rowDistribution(n, world, r) for r in range(world)]
            d0 = dists[0]
            assert d0.log2_block == 16 and (n >> d0.log2_block) >= world * 256
            counts = [d.n_local for d in dists]
            assert sum(counts) == n and max(counts) == d0.chunk
            assert (max(counts) - min(counts)) / max(counts) < 1e-3
            assert d0.chunk * world < 2**32 - 1
            rng = np.random.default_rng(n % 1000 + world)
            for d in dists[:: max(1, world // 2)]:
                for i in rng.integers(0, d.n_local, size=50):
                    g = ffi.lib().sped_dist_local_to_global(C.byref(d), int(i))
                    assert g < n and (g >> d.log2_block) % world == d.rank
                    assert ffi.lib().sped_dist_global_to_position(C.byref(d), g) == d.rank * d.chunk + int(i)


def test_config_defaults_and_parsing():
    # /root/reference/src/SpinED.hs:158-173
    spec = config.parseConfig(decks.load("heisenberg_chain_4"))
    assert spec.output == "exact_diagonalization_result.h5" and spec.number_vectors == 1
    assert spec.precision == 0.0 and spec.datatype == "float64"
    spec = config.parseConfig(decks.load("heisenberg_square_6x6"))
    assert spec.datatype == "float32" and spec.max_primme_basis_size == 20 and spec.max_primme_block_size == 4
    with pytest.raises(ffi.SpinEDException):
        config.parseDatatype("complex64")
    with pytest.raises(ffi.SpinEDException):
        config.parseConfig({"basis": {"number_spins": 2, "symmetries": []}})


def _jit_source(basis):
    need = C.c_uint64(0)
    ffi.checkStatus(ffi.lib().sped_selftest_jit_source(basis._ptr, None, 0, C.byref(need)))
    buf = C.create_string_buffer(need.value)
    ffi.checkStatus(ffi.lib().sped_selftest_jit_source(basis._ptr, buf, need.value, C.byref(need)))
    return buf.value.decode()


@pytest.mark.parametrize("name", ["heisenberg_chain_10", "heisenberg_square_4x4", "heisenberg_square_5x5",
                                  "heisenberg_pyrochlore_32", "heisenberg_square_6x6", "heisenberg_chain_42",
                                  "chain_50_hw2", "chain_64_hw2", "chain_40_hw3_k", "chain_8_k1_complex"])
def test_jit_generated_canonicalisation_matches_oracle(oracle, name, tmp_path):
    """The straight-line code handed to NVRTC, compiled for the host with g++ (funnel shift
    emulated), must canonicalise exactly like the oracle."""
    import subprocess

    cfg = extra_configs()[name] if name in extra_configs() else decks.load(name)
    ob, _ = oracle_problem(oracle, cfg)
    uc = product_problem(cfg)
    src = _jit_source(uc.cBasis)
    shim = r'''
#include <cstdint>
typedef unsigned long u64; typedef unsigned int u32;
#define __device__
#define __forceinline__ inline
static inline u32 __funnelshift_l(u32 lo, u32 hi, u32 s) { s &= 31u; return s ? (hi << s) | (lo >> (32u - s)) : hi; }
namespace sped { typedef ::u64 u64; typedef ::u32 u32; }
''' + src + r'''
extern "C" void canon(u64 n, const u64* x, u64* rep, int* phase) {
  for (u64 i = 0; i < n; ++i) sped::sped_jit_canonicalize(x[i], rep[i], phase[i]);
}
'''
    cpp = tmp_path / "jit_host.cpp"
    cpp.write_text(shim)
    so = tmp_path / "jit_host.so"
    subprocess.check_call(["/usr/bin/g++", "-O1", "-shared", "-fPIC", "-o", str(so), str(cpp)])
    L = C.CDLL(str(so))
    n, hw = ob.number_spins, ob.hamming_weight
    rng = np.random.default_rng(99)
    states = np.array([int(sum(1 << int(p) for p in rng.choice(n, size=hw, replace=False))) for _ in range(400)],
                      dtype=np.uint64)
    reps = np.zeros(len(states), dtype=np.uint64)
    phases = np.zeros(len(states), dtype=np.int32)
    L.canon(C.c_uint64(len(states)), C.c_void_p(states.ctypes.data), C.c_void_p(reps.ctypes.data), C.c_void_p(phases.ctypes.data))
    denom = np.lcm.reduce([2] + [oracle.periodicity(s["permutation"]) for s in cfg["basis"]["symmetries"]])
    for x, r, ph in zip(states, reps, phases):
        orep, ochi, onorm = ob.state_info(int(x))
        assert int(r) == orep
        if onorm > 0:
            assert abs(np.exp(2j * np.pi * ph / denom) - ochi) < 1e-12


def test_jit_kernel_compiles_for_sm_100a_without_a_gpu():
    uc = product_problem(decks.load("heisenberg_square_4x4"))
    out = C.c_uint64(0)
    rc = ffi.lib().sped_selftest_jit_compile(uc.cBasis._ptr, ffi.F64, 1, C.byref(out))
    if rc != 0:
        pytest.skip("NVRTC not available here: " + ffi.getErrorMessage(rc))
    assert out.value > 10000
