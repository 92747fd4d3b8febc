"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on
the same inputs.  Bar: representatives bit-exact; matvec <= 1e-12 relative L2 (f64/c128 storage;
f32/c64 storage is compared at 5e-6, the rounding of the stored vectors); eigenvalues <= 1e-10
relative; residuals <= 1e-8 (relative to |E0|)."""
import numpy as np
import pytest

from helpers import SMALL_DECKS, extra_configs, oracle_problem, product_problem, splitmix_vector
from spin_ed_b200 import decks, ffi

pytestmark = pytest.mark.gpu

ALL_SMALL = {**{n: None for n in SMALL_DECKS}, **extra_configs()}


def _cfg(name):
    return decks.load(name) if ALL_SMALL.get(name) is None and name not in extra_configs() else extra_configs()[name]


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def test_device_present_and_library_loaded():
    assert ffi.deviceCount() >= 1
    assert b"sm_100a" in ffi.lib().sped_version()


@pytest.mark.parametrize("name", sorted(ALL_SMALL))
def test_representatives_bit_exact_small(oracle, name):
    cfg = _cfg(name)
    ob, _ = oracle_problem(oracle, cfg)
    ob.build()
    uc = product_problem(cfg)
    ffi.buildBasis(uc.cBasis)
    assert ffi.getNumberStates(uc.cBasis) == ob.number_states
    got = ffi.basisGetStates(uc.cBasis)
    assert got.dtype == np.uint64 and np.array_equal(got, ob.states)
    assert np.array_equal(ffi.basisNorms(uc.cBasis), ob.norms)


def test_known_answer_readme_ring():
    # /root/reference/README.md:37-95 -- 4-spin ring: 16 representatives, E0 = -8
    uc = product_problem(decks.load("heisenberg_chain_4"))
    ffi.buildBasis(uc.cBasis)
    assert np.array_equal(ffi.basisGetStates(uc.cBasis), np.arange(16, dtype=np.uint64))
    ev, vecs, rn = ffi.eigh(uc.cHamiltonian.operatorObject, np.float64, 1)
    assert abs(ev[0] + 8.0) < 1e-10 and rn[0] < 1e-8
    assert vecs.shape == (16, 1)


def test_known_answer_chain_10_representatives():
    uc = product_problem(decks.load("heisenberg_chain_10"))
    ffi.buildBasis(uc.cBasis)
    assert list(ffi.basisGetStates(uc.cBasis)) == [31, 47, 55, 87, 91, 93, 103, 107, 155, 171, 173, 179, 341]


@pytest.mark.parametrize("name", sorted(ALL_SMALL))
@pytest.mark.parametrize("block", [1, 3])
def test_matvec_matches_oracle_small(oracle, name, block):
    cfg = _cfg(name)
    ob, terms = oracle_problem(oracle, cfg)
    ob.build()
    oop = oracle.Operator(ob, terms)
    uc = product_problem(cfg)
    ffi.buildBasis(uc.cBasis)
    op = uc.cHamiltonian.operatorObject
    assert ffi.isOperatorReal(op) == oop.is_real
    n = ob.number_states
    dtypes = [np.float64, np.float32] if oop.is_real else []
    dtypes += [np.complex128, np.complex64]
    for dt in dtypes:
        x = np.asfortranarray(np.stack([splitmix_vector(n, 0x5EED0001 + c, dt) for c in range(block)], axis=1))
        want = oop.matmat(x.astype(np.complex128 if np.dtype(dt).kind == "c" else np.float64))
        got = ffi.apply(op, x)
        tol = 1e-12 if np.dtype(dt).itemsize // (2 if np.dtype(dt).kind == "c" else 1) == 8 else 5e-6
        assert got.dtype == np.dtype(dt)
        assert rel(got, want) < tol, (name, dt, rel(got, want))
    rows, E = ffi.operatorCountElements(op)
    assert rows == n and E == oop.count_offdiag()
    if not oop.is_real:
        with pytest.raises(ffi.LatticeSymmetriesException) as e:
            ffi.apply(op, np.zeros(n, dtype=np.float64))
        assert e.value.eCode == 18


def test_matvec_strided_blocks_and_dimension_mismatch(oracle):
    cfg = decks.load("heisenberg_square_4x4")
    ob, terms = oracle_problem(oracle, cfg)
    ob.build()
    oop = oracle.Operator(ob, terms)
    uc = product_problem(cfg)
    ffi.buildBasis(uc.cBasis)
    op = uc.cHamiltonian.operatorObject
    n = ob.number_states
    big = np.zeros((n + 5, 4), order="F")
    big[:n, :] = np.stack([splitmix_vector(n, 7 + c) for c in range(4)], axis=1)
    x = big[:n, :]  # column stride n + 5
    y = np.zeros((n, 4), order="F")
    ffi.checkStatus(ffi.lib().ls_operator_matmat(op._ptr, ffi.F64, n, 4, x.ctypes.data, n + 5, y.ctypes.data, n))
    assert rel(y, oop.matmat(np.asfortranarray(x))) < 1e-12
    with pytest.raises(ffi.LatticeSymmetriesException) as e:
        ffi.apply(op, np.zeros(n + 1))
    assert e.value.eCode == 19


@pytest.mark.parametrize("name", ["heisenberg_chain_10", "heisenberg_square_4x4", "chain_8_k1_complex", "ring_4site_nosym"])
def test_expectation_matches_oracle(oracle, name):
    cfg = _cfg(name)
    ob, terms = oracle_problem(oracle, cfg)
    ob.build()
    oop = oracle.Operator(ob, terms)
    uc = product_problem(cfg)
    ffi.buildBasis(uc.cBasis)
    n = ob.number_states
    dt = np.float64 if oop.is_real else np.complex128
    x = np.asfortranarray(np.stack([splitmix_vector(n, 3 + c, dt) for c in range(2)], axis=1))
    got = ffi.expectation(uc.cHamiltonian.operatorObject, x)
    want = oop.expectation(x)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)


def test_build_unsafe_adopts_representatives(oracle):
    cfg = decks.load("heisenberg_square_4x4")
    ob, terms = oracle_problem(oracle, cfg)
    ob.build()
    uc = product_problem(cfg)
    ffi.buildBasis(uc.cBasis, ob.states)  # resume path, /root/reference/src/SpinED.hs:319-331
    assert ffi.getNumberStates(uc.cBasis) == 107
    x = splitmix_vector(107)
    assert rel(ffi.apply(uc.cHamiltonian.operatorObject, x), oracle.Operator(ob, terms).matmat(x)) < 1e-12
    bad = ob.states.copy()
    bad[3] = bad[3] ^ np.uint64(3)  # no longer an orbit minimum (or unsorted)
    uc2 = product_problem(cfg)
    with pytest.raises(ffi.LatticeSymmetriesException):
        ffi.buildBasis(uc2.cBasis, bad)


def test_operator_created_before_build_sees_later_build(oracle):
    # /root/reference/src/SpinED.hs:243-248 creates operators before app/Main.hs:22 builds the basis
    cfg = decks.load("heisenberg_chain_10")
    uc = product_problem(cfg)
    op = uc.cHamiltonian.operatorObject
    with pytest.raises(ffi.LatticeSymmetriesException) as e:
        ffi.apply(op, np.zeros(13))
    assert e.value.eCode == 14
    ffi.buildBasis(uc.cBasis)
    ob, terms = oracle_problem(oracle, cfg)
    ob.build()
    x = splitmix_vector(13)
    assert rel(ffi.apply(op, x), oracle.Operator(ob, terms).matmat(x)) < 1e-12


def test_handles_survive_any_destroy_order(oracle):
    import gc

    cfg = decks.load("heisenberg_chain_10")
    uc = product_problem(cfg)
    ffi.buildBasis(uc.cBasis)
    states = ffi.basisGetStates(uc.cBasis)
    op = uc.cHamiltonian.operatorObject
    del uc
    gc.collect()
    assert list(states[:3]) == [31, 47, 55]
    ob, terms = oracle_problem(oracle, cfg)
    ob.build()
    x = splitmix_vector(13)
    assert rel(ffi.apply(op, x), oracle.Operator(ob, terms).matmat(x)) < 1e-12


@pytest.mark.parametrize("name,k", [("heisenberg_chain_10", 1), ("heisenberg_kagome_12", 4), ("heisenberg_square_4x4", 2),
                                    ("heisenberg_triangular_19", 2), ("chain_8_k1_complex", 2), ("chain_12_full_sym", 3)])
def test_eigenpairs_match_oracle_dense(oracle, name, k):
    cfg = _cfg(name)
    ob, terms = oracle_problem(oracle, cfg)
    ob.build()
    oop = oracle.Operator(ob, terms)
    n = ob.number_states
    if n <= 1500:
        want = np.linalg.eigvalsh(oop.to_dense())[:k]
    else:
        import scipy.sparse.linalg as sla

        dt = np.float64 if oop.is_real else np.complex128
        A = sla.LinearOperator((n, n), matvec=lambda v: oop.matmat(np.ascontiguousarray(v, dtype=dt)), dtype=dt)
        want = np.sort(sla.eigsh(A, k=k + 2, which="SA", tol=1e-12)[0])[:k]
    uc = product_problem(cfg)
    ffi.buildBasis(uc.cBasis)
    op = uc.cHamiltonian.operatorObject
    dt = np.float64 if ffi.isOperatorReal(op) else np.complex128
    ev, vecs, rn = ffi.eigh(op, dt, k)
    scale = max(1.0, abs(want[0]))
    assert np.all(np.abs(ev - want) <= 1e-10 * scale), (ev, want)
    assert np.all(rn <= 1e-8 * scale)
    # eigenvectors: residual of what was returned, checked with the oracle's operator
    hv = oop.matmat(np.asfortranarray(vecs))
    for i in range(k):
        assert np.linalg.norm(hv[:, i] - ev[i] * vecs[:, i]) <= 1e-8 * scale
        assert abs(np.linalg.norm(vecs[:, i]) - 1) < 1e-10
    # expectation of H in the eigenvectors reproduces the eigenvalues (observables path)
    ex = ffi.expectation(op, vecs)
    assert np.allclose(ex.real, ev, atol=1e-9 * scale) and np.allclose(ex.imag, 0, atol=1e-9)


def test_eigh_small_basis_restarts_like_the_40_spin_decks(oracle):
    # chain_40/42 run with max_primme_basis_size 3/4: exercise restarts on a small symmetric chain
    cfg = decks.chain(16, 8, 1, (0, 0))
    ob, terms = oracle_problem(oracle, cfg)
    ob.build()
    want = np.linalg.eigvalsh(oracle.Operator(ob, terms).to_dense())[0]
    uc = product_problem(cfg)
    ffi.buildBasis(uc.cBasis)
    op = uc.cHamiltonian.operatorObject
    ev, _, rn = ffi.eigh(op, np.float64, 1, maxBasisSize=3)
    st = ffi.eighLastStats(op)
    assert abs(ev[0] - want) <= 1e-10 * abs(want) and rn[0] <= 1e-8 * abs(want)
    assert st["restarts"] > 0 and st["matvecs"] >= st["iterations"]


@pytest.mark.parametrize("mmax", [2, 3, 4, 5, 8])
def test_eigh_single_pair_small_basis_fused_restart(oracle, mmax):
    """One wanted pair and a basis of <= 8 vectors: restart + residual are one fused pass and the
    residual goes straight into the next basis column (eigh.cu, `single`).  Same eigenvalue as the
    dense spectrum, residual within tolerance, eigenvector an eigenvector -- real and complex."""
    for cfg, dt in ((decks.chain(16, 8, 1, (0, 0)), np.float64), (extra_configs()["chain_8_k1_complex"], np.complex128),
                    (decks.chain(14, 7, None, (0, 0)), np.complex128)):
        ob, terms = oracle_problem(oracle, cfg)
        ob.build()
        dense = oracle.Operator(ob, terms).to_dense()
        want = np.linalg.eigvalsh(dense)[0]
        uc = product_problem(cfg)
        ffi.buildBasis(uc.cBasis)
        op = uc.cHamiltonian.operatorObject
        ev, vecs, rn = ffi.eigh(op, dt, 1, maxBasisSize=mmax)
        st = ffi.eighLastStats(op)
        scale = np.abs(np.linalg.eigvalsh(dense)).max()
        assert abs(ev[0] - want) <= 1e-10 * scale, (mmax, dt, ev[0], want)
        assert rn[0] <= 1e-8 * scale
        v = vecs[:, 0] if vecs.ndim == 2 else vecs
        assert abs(np.linalg.norm(v) - 1) < 1e-10
        assert np.linalg.norm(dense @ v - ev[0] * v) <= 1e-8 * scale
        if ob.number_states > 3 * mmax:
            assert st["restarts"] > 0


def test_eigh_single_pair_float32_storage_small_basis():
    cfg = decks.load("heisenberg_square_4x4")
    uc = product_problem(cfg)
    ffi.buildBasis(uc.cBasis)
    ev, vecs, rn = ffi.eigh(uc.cHamiltonian.operatorObject, np.float32, 1, maxBasisSize=3)
    assert vecs.dtype == np.float32
    assert abs(ev[0] + 44.9139328337) < 2e-3


def test_eigh_float32_storage(oracle):
    cfg = decks.load("heisenberg_square_4x4")  # the deck asks for float32 (SpinED.hs:344-352)
    uc = product_problem(cfg)
    ffi.buildBasis(uc.cBasis)
    ev, vecs, rn = ffi.eigh(uc.cHamiltonian.operatorObject, np.float32, 2, maxBasisSize=20, maxBlockSize=4)
    assert vecs.dtype == np.float32
    assert abs(ev[0] + 44.9139328337) < 2e-3  # storage-precision bound, float64 parity is tested above


def test_monitor_callback_and_abort():
    uc = product_problem(decks.load("heisenberg_kagome_12"))
    ffi.buildBasis(uc.cBasis)
    seen = []
    ffi.eigh(uc.cHamiltonian.operatorObject, np.float64, 1, monitor=lambda info: seen.append(info) or False)
    assert seen and seen[-1]["number_converged"] == 1 and seen[0]["iteration"] == 0


@pytest.mark.parametrize("name", ["heisenberg_square_5x5", "heisenberg_chain_24", "xxz_triangular_19", "heisenberg_pyrochlore_32"])
def test_mid_size_decks_bit_exact_and_matvec(oracle, name):
    cfg = decks.load(name)
    ob, terms = oracle_problem(oracle, cfg)
    ob.build()
    oop = oracle.Operator(ob, terms)
    uc = product_problem(cfg)
    ffi.buildBasis(uc.cBasis)
    assert np.array_equal(ffi.basisGetStates(uc.cBasis), ob.states)
    n = ob.number_states
    dt = np.float64 if oop.is_real else np.complex128
    x = splitmix_vector(n, 0x5EED0001, dt)
    x /= np.linalg.norm(x)
    got = ffi.apply(uc.cHamiltonian.operatorObject, x)
    want, E = oop.matmat(x, count=True)
    assert rel(got, want) < 1e-12
    assert ffi.operatorCountElements(uc.cHamiltonian.operatorObject) == (n, E)


def test_full_size_6x6_properties():
    """BASELINE full size (no oracle run: minutes of CPU): size-independent properties."""
    uc = product_problem(decks.load("heisenberg_square_6x6"))
    ffi.buildBasis(uc.cBasis)
    n = ffi.getNumberStates(uc.cBasis)
    assert n == 15804956  # Burnside / literature value
    reps = ffi.basisGetStates(uc.cBasis)
    assert np.all(reps[1:] > reps[:-1])
    pop = np.zeros(n, dtype=np.uint8)
    for b in range(36):
        pop += ((reps >> np.uint64(b)) & np.uint64(1)).astype(np.uint8)
    assert np.all(pop == 18)
    # every representative is its own representative with non-zero norm (idempotence of canonicalise)
    sample = reps[:: max(1, n // 4096)]
    r2, chi, norms = ffi.basisStateInfo(uc.cBasis, sample)
    assert np.array_equal(r2, sample) and np.all(norms > 0) and np.allclose(chi, 1)
    # Hermiticity: <u, H v> == <H u, v>
    op = uc.cHamiltonian.operatorObject
    u = splitmix_vector(n, 11)
    v = splitmix_vector(n, 12)
    hu, hv = ffi.apply(op, u), ffi.apply(op, v)
    assert abs(np.dot(u, hv) - np.dot(hu, v)) <= 1e-10 * abs(np.dot(u, hv))
    # linearity
    w = ffi.apply(op, 2.0 * u - 3.0 * v)
    assert rel(w, 2.0 * hu - 3.0 * hv) < 1e-12


@pytest.mark.parametrize("name", sorted(ALL_SMALL) + ["heisenberg_square_5x5", "heisenberg_chain_24"])
def test_operator_cache_agrees_with_matrix_free(oracle, name):
    """The HBM-resident operator cache holds the same elements as the matrix-free kernel finds; it
    sums the default-coefficient elements before the coded ones (and multiplies their common
    coefficient once), so the two agree to rounding, for every storage type and block width, and
    the cached path is reproducible bit for bit."""
    cfg = _cfg(name) if name in ALL_SMALL else decks.load(name)
    uc = product_problem(cfg)
    ffi.buildBasis(uc.cBasis)
    op = uc.cHamiltonian.operatorObject
    n = ffi.getNumberStates(uc.cBasis)
    dtypes = ([np.float64, np.float32] if ffi.isOperatorReal(op) else []) + [np.complex128, np.complex64]
    for dt in dtypes:
        for block in (1, 5):
            x = np.asfortranarray(np.stack([splitmix_vector(n, 21 + c, dt) for c in range(block)], axis=1))
            ffi.operatorSetCache(op, 0)
            free = ffi.apply(op, x)
            assert not ffi.operatorCacheInfo(op)["ready"]
            ffi.operatorSetCache(op, 1)
            cached = ffi.apply(op, x)
            info = ffi.operatorCacheInfo(op)
            assert info["ready"] and info["bytes"] > 0
            tol = 1e-13 if np.dtype(dt).itemsize >= 8 and np.dtype(dt) != np.complex64 else 1e-6
            assert np.linalg.norm(free - cached) <= tol * np.linalg.norm(free), (name, dt, block)
            again = ffi.apply(op, x)
            assert again.tobytes() == cached.tobytes()


@pytest.mark.parametrize("name", ["heisenberg_chain_10", "chain_12_pi", "chain_8_k1_complex", "heisenberg_triangular_19",
                                  "heisenberg_square_5x5"])
def test_operator_cache_fill_in_row_chunks(name, monkeypatch):
    """The cache fill runs in row chunks (its one-code-per-slot temporary covers one chunk) and keeps
    only the codes of the coded elements.  Chunks of a few hundred bytes force many of them on small
    decks: the cache must be the same size and give the same bits as the one filled in one piece."""
    cfg = _cfg(name) if name in ALL_SMALL else decks.load(name)
    results = []
    for chunk in ("", "300", "1"):
        monkeypatch.setenv("SPED_FILL_CHUNK_BYTES", chunk)
        uc = product_problem(cfg)
        ffi.buildBasis(uc.cBasis)
        op = uc.cHamiltonian.operatorObject
        n = ffi.getNumberStates(uc.cBasis)
        dt = np.float64 if ffi.isOperatorReal(op) else np.complex128
        x = np.asfortranarray(np.stack([splitmix_vector(n, 77 + c, dt) for c in range(3)], axis=1))
        ffi.operatorSetCache(op, 1)
        y = ffi.apply(op, x)
        info = ffi.operatorCacheInfo(op)
        assert info["ready"]
        results.append((y.tobytes(), info["bytes"]))
    assert results[0] == results[1] == results[2]


def test_cache_filled_by_either_kernel_is_the_same(monkeypatch):
    """NVRTC runs in the background from ls_build on; a cache fill that comes too early for it uses the
    interpreted kernel.  Whichever kernel fills the cache, the cached product has the same bits."""
    cfg = decks.load("heisenberg_square_5x5")
    outs = []
    for env in ({"SPED_JIT": "0"}, {"SPED_JIT": "1", "SPED_JIT_WAIT": "1"}, {"SPED_JIT": "1", "SPED_CACHE_DIR": ""}):
        for k in ("SPED_JIT", "SPED_JIT_WAIT", "SPED_CACHE_DIR"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        uc = product_problem(cfg)
        ffi.buildBasis(uc.cBasis)
        op = uc.cHamiltonian.operatorObject
        ffi.operatorSetCache(op, 1)
        x = splitmix_vector(ffi.getNumberStates(uc.cBasis), 9, np.complex128)  # complex characters (k != 0)
        outs.append(ffi.apply(op, x).tobytes())
        assert ffi.operatorCacheInfo(op)["ready"]
    assert outs[0] == outs[1] == outs[2]


def test_two_byte_coefficient_codes(oracle, monkeypatch):
    """More than 256 distinct (value, phase, stabiliser) triples: the operator cache stores u16 codes.
    No shipped deck needs them; a 20-site J1-J2-J3 chain in momentum sector 1 does (360 codes).  Cached
    (one chunk and many chunks), matrix-free and oracle results must agree."""
    from test_emulation import _wide_code_config

    cfg = _wide_code_config()
    ob, terms = oracle_problem(oracle, cfg)
    ob.build()
    oop = oracle.Operator(ob, terms)
    n = ob.number_states
    x = np.asfortranarray(np.stack([splitmix_vector(n, 31 + c, np.complex128) for c in range(3)], axis=1))
    want = oop.matmat(x)
    outs = []
    for chunk in ("", "500"):
        monkeypatch.setenv("SPED_FILL_CHUNK_BYTES", chunk)
        uc = product_problem(cfg)
        ffi.buildBasis(uc.cBasis)
        assert np.array_equal(ffi.basisGetStates(uc.cBasis), ob.states)
        op = uc.cHamiltonian.operatorObject
        ffi.operatorSetCache(op, 0)
        free = ffi.apply(op, x)
        ffi.operatorSetCache(op, 1)
        cached = ffi.apply(op, x)
        assert ffi.operatorCacheInfo(op)["ready"]
        for got in (free, cached):
            assert np.linalg.norm(got - want) <= 1e-12 * np.linalg.norm(want)
        outs.append(cached.tobytes())
    assert outs[0] == outs[1]


def test_jit_and_interpreted_kernels_agree_bitwise(monkeypatch):
    cfg = decks.load("heisenberg_square_5x5")
    n_expected = 208012
    outs = []
    monkeypatch.setenv("SPED_JIT_WAIT", "1")  # the specialised kernel itself must run, not its interpreted stand-in
    for jit in ("1", "0"):
        monkeypatch.setenv("SPED_JIT", jit)
        uc = product_problem(cfg)
        ffi.buildBasis(uc.cBasis)
        op = uc.cHamiltonian.operatorObject
        ffi.operatorSetCache(op, 0)
        assert ffi.getNumberStates(uc.cBasis) == n_expected
        x = splitmix_vector(n_expected, 5, np.complex128)
        outs.append(ffi.apply(op, x))
    assert outs[0].tobytes() == outs[1].tobytes()


def test_gpu_matches_committed_golden_vectors():
    """The CUDA path against tests/golden/small_sectors.json (made by an independent numpy
    construction, see tests/golden/make_golden.py)."""
    import json
    import os

    golden = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "small_sectors.json")))
    for name, g in sorted(golden.items()):
        uc = product_problem(_cfg(name))
        ffi.buildBasis(uc.cBasis)
        assert [int(v) for v in ffi.basisGetStates(uc.cBasis)] == g["representatives"], name
        assert np.allclose(ffi.basisNorms(uc.cBasis), g["norms"], rtol=0, atol=1e-15)
        x = np.array([complex(*v) for v in g["x"]])
        y = np.array([complex(*v) for v in g["y"]])
        got = ffi.apply(uc.cHamiltonian.operatorObject, x if g["complex"] else x.real.copy())
        assert rel(got, y) < 1e-12, name
        k = min(3, len(g["eigenvalues"]))
        ev, _, rn = ffi.eigh(uc.cHamiltonian.operatorObject, np.complex128 if g["complex"] else np.float64, k)
        scale = max(1.0, abs(g["eigenvalues"][0]))
        assert np.all(np.abs(ev - np.array(g["eigenvalues"][:k])) <= 1e-10 * scale), (name, ev, g["eigenvalues"][:k])


def test_degenerate_levels_block_solver(oracle):
    """Degenerate multiplets: a chain without any symmetry restriction has SU(2) multiplets spread
    over magnetisation sectors; the block solver must return each level with its multiplicity."""
    cfg = {"basis": {"number_spins": 10, "symmetries": []}, "hamiltonian": decks.chain(10)["hamiltonian"], "observables": []}
    ob, terms = oracle_problem(oracle, cfg)
    ob.build()
    want = np.linalg.eigvalsh(oracle.Operator(ob, terms).to_dense())[:7]
    uc = product_problem(cfg)
    ffi.buildBasis(uc.cBasis)
    ev, vecs, rn = ffi.eigh(uc.cHamiltonian.operatorObject, np.float64, 7, maxBlockSize=4)
    assert np.all(np.abs(ev - want) <= 1e-9 * abs(want[0])), (ev, want)
    assert np.allclose(vecs.T @ vecs, np.eye(7), atol=1e-9)


def test_xxz_triangular_19_eigenvectors_are_orthonormal_eigenvectors(oracle):
    """configs[1]: six lowest states (number_vectors 6, block 6, precision 1e-6).  Multiplicities
    are certified by the vectors themselves: orthonormal and each an eigenvector (oracle matvec)."""
    cfg = decks.load("xxz_triangular_19")
    ob, terms = oracle_problem(oracle, cfg)
    ob.build()
    oop = oracle.Operator(ob, terms)
    uc = product_problem(cfg)
    ffi.buildBasis(uc.cBasis)
    ev, vecs, rn = ffi.eigh(uc.cHamiltonian.operatorObject, np.float64, 6, 1.0e-6, 0, 6, 0)
    assert np.allclose(vecs.T @ vecs, np.eye(6), atol=1e-8)
    hv = oop.matmat(np.asfortranarray(vecs))
    a_norm = 20.0
    for i in range(6):
        assert np.linalg.norm(hv[:, i] - ev[i] * vecs[:, i]) <= 2e-6 * a_norm, (i, ev)
    assert abs(ev[0] + 8.6351360078) < 1e-6
    print("xxz_triangular_19 lowest six:", ev)


@pytest.mark.parametrize("n", [16, 24])
def test_ground_state_energy_agrees_between_sector_descriptions(n):
    """Size-independent check used at 36-42 spins (tools/sector_cross_check.py): the fully symmetric
    sector (translations x parity x spin inversion) and the larger translations-only sector share
    nothing but the physics -- different group, representatives, norms, basis size -- yet hold the
    same ground state, so E0 must agree to 1e-10 relative."""
    full = decks.chain(n, n // 2, 1, (0, 0))
    larger = decks.chain(n, n // 2, 1, (0, 0))
    larger["basis"]["symmetries"] = larger["basis"]["symmetries"][:1]
    e0, dims = [], []
    for cfg in (full, larger):
        uc = product_problem(cfg)
        ffi.buildBasis(uc.cBasis)
        dims.append(ffi.getNumberStates(uc.cBasis))
        ev, _, rn = ffi.eigh(uc.cHamiltonian.operatorObject, np.float64, 1, want_vectors=False)
        assert rn[0] <= 1e-8 * max(1.0, abs(ev[0]))
        e0.append(ev[0])
    assert dims[1] > 1.5 * dims[0]
    assert abs(e0[0] - e0[1]) <= 1e-10 * abs(e0[0]), (e0, dims)


@pytest.mark.parametrize("n", [16, 20, 24])
def test_ground_state_energy_equals_bethe_ansatz(n):
    """The exact Bethe-ansatz energy of the Heisenberg ring (oracle/bethe.py: no basis, no symmetry
    group, no sparse product, no eigensolver in common with the product) against sped_eigh in the fully
    symmetric sector that holds the ground state for even n/2."""
    from oracle import bethe

    uc = product_problem(decks.chain(n, n // 2, 1, (0, 0)))
    ffi.buildBasis(uc.cBasis)
    ev, _, rn = ffi.eigh(uc.cHamiltonian.operatorObject, np.float64, 1)
    exact = bethe.sigma_sigma_ring_energy(n)
    assert abs(ev[0] - exact) <= 1e-10 * abs(exact), (n, ev[0], exact)
    assert rn[0] <= 1e-8 * abs(exact)


@pytest.mark.parametrize("n", [16, 20])
def test_xx_ring_ground_state_equals_free_fermions(n):
    """XX ring (two-site matrix without a diagonal part): sped_eigh in the symmetric sector against the
    exact free-fermion energy (oracle/bethe.py)."""
    from oracle import bethe
    from test_oracle import xx_chain

    uc = product_problem(xx_chain(n, 1, (0, 0)))
    ffi.buildBasis(uc.cBasis)
    ev, _, rn = ffi.eigh(uc.cHamiltonian.operatorObject, np.float64, 1)
    exact = bethe.xx_ring_energy(n)
    assert abs(ev[0] - exact) <= 1e-10 * abs(exact), (n, ev[0], exact)
    assert rn[0] <= 1e-8 * abs(exact)
