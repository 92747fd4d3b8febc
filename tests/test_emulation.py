"""The device sources of the matvec kernels, compiled by the HOST compiler with the CUDA built-ins
shimmed (csrc/emul.cpp, one thread), must reproduce the oracle on small decks: the matrix-free row
routine, the cache fill (source classes, default-coefficient / coded sub-classes, two-ended slots)
and the streaming kernel -- all source classes in one pass and class by class -- for 1 to 8 ranks.
No GPU involved: this checks the slot arithmetic and summation logic, not the hardware path (the
-m gpu tests do that)."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import SMALL_DECKS, extra_configs, oracle_problem, product_problem, splitmix_vector
from spin_ed_b200 import decks, ffi

NAMES = SMALL_DECKS + ["chain_12_full_sym", "chain_12_pi", "chain_8_k1_complex", "chain_9_k2_nohw", "chain_10_inv_only",
                       "chain_10_inv_nohw", "chain_8_chiral_3site", "ring_4site_nosym", "chain_40_hw3_k", "xx_chain_12_sym"]


def _emulate(op, reps, stab, world, rank, x, ncols=1):
    n = len(reps)
    rd = ffi.rowDistribution(n, world, rank)
    n_local = int(rd.n_local)
    dt = x.dtype
    outs = [np.full(max(n_local, 1), 7.0, dtype=dt) for _ in range(3)]
    block = np.full((max(n_local, 1), max(ncols, 1)), 7.0, dtype=dt, order="F")
    stats = (C.c_uint64 * 6)()
    ffi.checkStatus(ffi.emulLib().sped_selftest_emulate_matvec(
        op._ptr, n, reps.ctypes.data, stab.ctypes.data, world, rank, ffi.DTYPE_TAGS[np.dtype(dt)], x.ctypes.data,
        outs[0].ctypes.data, outs[1].ctypes.data, outs[2].ctypes.data, stats, ncols, block.ctypes.data))
    rows = rd.local_rows().astype(np.int64)
    if ncols > 1:
        return rows, [o[:n_local] for o in outs], [int(v) for v in stats], block[:n_local]
    return rows, [o[:n_local] for o in outs], [int(v) for v in stats]


def _wide_code_config():
    """20-site J1-J2-J3 chain, momentum sector 1 of the translations: 3 distinct off-diagonal values x 20
    phases x 6 stabiliser sizes = 360 coefficient codes -- more than a byte holds, so the cache stores
    u16 codes (no shipped deck does)."""
    n = 20

    def heis(j):
        return [[j, 0, 0, 0], [0, -j, 2 * j, 0], [0, 2 * j, -j, 0], [0, 0, 0, j]]

    return {"basis": {"number_spins": n, "hamming_weight": n // 2,
                      "symmetries": [{"permutation": [(i + 1) % n for i in range(n)], "sector": 1}]},
            "hamiltonian": {"name": "J1-J2-J3", "terms": [
                {"matrix": heis(1.0), "sites": [[i, (i + 1) % n] for i in range(n)]},
                {"matrix": heis(0.4), "sites": [[i, (i + 2) % n] for i in range(n)]},
                {"matrix": heis(0.15), "sites": [[i, (i + 3) % n] for i in range(n)]}]},
            "observables": []}


def test_emulated_kernels_with_two_byte_codes(oracle, monkeypatch):
    monkeypatch.setenv("SPED_REMOTE_GROUPS", "2")
    cfg = _wide_code_config()
    ob, terms = oracle_problem(oracle, cfg)
    ob.build()
    oop = oracle.Operator(ob, terms)
    n = ob.number_states
    reps = np.ascontiguousarray(ob.states, dtype=np.uint64)
    stab = np.ascontiguousarray(np.rint(ob.norms ** 2 * ob.group_size), dtype=np.uint16)
    op = product_problem(cfg).cHamiltonian.operatorObject
    x = np.ascontiguousarray(splitmix_vector(n, 0x5EED0001, np.complex128))
    want = oop.matmat(x)
    scale = np.linalg.norm(want)
    for world in (1, 4):
        for rank in range(world):
            rows, (free, allc, phased), stats = _emulate(op, reps, stab, world, rank, x)
            assert stats[5] == 2, "this configuration is meant to need u16 codes"
            assert stats[1] > stats[2] > 0  # coded and default elements both occur
            for got in (free, allc, phased):
                assert np.linalg.norm(got - want[rows]) <= 1e-12 * scale
    xb = np.asfortranarray(np.stack([np.roll(x, -c) for c in range(3)], axis=1))
    wantb = oop.matmat(xb)
    rows, _, _, blk = _emulate(op, reps, stab, 1, 0, x, 3)
    assert np.linalg.norm(blk - wantb[rows]) <= 1e-12 * np.linalg.norm(wantb)


@pytest.mark.parametrize("groups", ["1", "2"])
@pytest.mark.parametrize("name", NAMES)
def test_emulated_kernels_match_oracle(oracle, name, groups, monkeypatch):
    # exchange rounds for > 2 ranks: shards this small would take one round (all-gather); both the
    # two-class and the three-class layout are forced in turn
    monkeypatch.setenv("SPED_REMOTE_GROUPS", groups)
    ONE_ROUND = groups == "1"
    cfg = extra_configs()[name] if name in extra_configs() else decks.load(name)
    ob, terms = oracle_problem(oracle, cfg)
    ob.build()
    oop = oracle.Operator(ob, terms)
    n = ob.number_states
    reps = np.ascontiguousarray(ob.states, dtype=np.uint64)
    order = ob.group_size * (2 if cfg["basis"].get("spin_inversion") else 1)
    stab = np.ascontiguousarray(np.rint(ob.norms ** 2 * order), dtype=np.uint16)
    assert np.all(stab >= 1)
    uc = product_problem(cfg)
    op = uc.cHamiltonian.operatorObject
    dt = np.float64 if oop.is_real else np.complex128
    x = np.ascontiguousarray(splitmix_vector(n, 0x5EED0001, dt))
    want = oop.matmat(x)
    scale = np.linalg.norm(want)
    total = oop.count_offdiag()
    for world in (1, 2, 3, 4, 8):
        elements = 0
        for rank in range(world):
            rows, (free, allc, phased), stats = _emulate(op, reps, stab, world, rank, x)
            elements += stats[1]
            rounds = 0 if world == 1 else 1 if (world == 2 or ONE_ROUND) else 2
            assert stats[3] == 1 + rounds
            assert stats[0] >= stats[1] >= stats[2]
            # compact code stream: room for every coded element, never more than one entry per slot,
            # and nothing at all when every element carries the default coefficient
            assert stats[1] - stats[2] <= stats[4] <= stats[0]
            assert (stats[4] == 0) == (stats[1] == stats[2])
            for what, got in (("matrix-free", free), ("cached, one pass", allc), ("cached, class by class", phased)):
                err = np.linalg.norm(got - want[rows])
                assert err <= 1e-12 * scale, (name, world, rank, what, err / scale)
            # the one-pass and the class-by-class streaming results are the same sums in the same order
            assert np.array_equal(allc, phased), (name, world, rank)
        assert elements == total, (name, world, elements, total)
    # block kernel (interleaved columns): 2, 3 and 4 columns, one and several ranks
    for ncols in (2, 3, 4):
        xb = np.asfortranarray(np.stack([np.roll(x, -c) for c in range(ncols)], axis=1))
        wantb = oop.matmat(xb)
        for world in (1, 3):
            for rank in range(world):
                rows, _, _, blk = _emulate(op, reps, stab, world, rank, x, ncols)
                if len(rows):
                    err = np.linalg.norm(blk - wantb[rows])
                    assert err <= 1e-12 * np.linalg.norm(wantb), (name, ncols, world, rank, err)


@pytest.mark.parametrize("dt", [np.float64, np.complex128, np.float32, np.complex64])
@pytest.mark.parametrize("m,p", [(2, 1), (3, 2), (4, 2), (4, 3), (5, 2), (8, 7)])
def test_emulated_fused_restart_residual_kernels(dt, m, p):
    """eigh_kernels.cuh on the host: the fused restart + residual pass of the single-pair eigensolver
    iteration, the axpy + norm pass and the relative scaling must equal the plain linear algebra they
    stand for: V' = V C, W' = W C, r = (W - theta V) C[:, 0] in column p of V, |r|^2 and V'^H r, then
    t = r - V' (V'^H r), |t|^2, t / |t|."""
    rng = np.random.default_rng(1000 * m + 10 * p + np.dtype(dt).itemsize)
    n = 777
    cplx = np.dtype(dt).kind == "c"
    wide = np.complex128 if cplx else np.float64

    def rand(*shape):
        a = rng.standard_normal(shape)
        return a + 1j * rng.standard_normal(shape) if cplx else a

    V = np.linalg.qr(rand(n, m))[0]
    A = rand(n, n)
    A = A + A.conj().T
    W = A @ V
    H = V.conj().T @ W
    theta = float(np.linalg.eigvalsh(0.5 * (H + H.conj().T))[0])
    C = np.linalg.qr(rand(m, p))[0].astype(np.complex128)
    Vs = np.asfortranarray(V.astype(dt))
    Ws = np.asfortranarray(W.astype(dt))
    V0, W0 = Vs.astype(wide), Ws.astype(wide)  # what the kernel reads (storage precision)
    Cw = C if cplx else C.real.astype(np.float64)
    if not cplx:
        C = C.real.astype(np.complex128)
    c_pairs = np.ascontiguousarray(np.stack([C.real, C.imag], axis=-1).reshape(-1))
    out = np.zeros(2 * (1 + p) + 3)
    ffi.checkStatus(ffi.emulLib().sped_selftest_emulate_restart(
        ffi.DTYPE_TAGS[np.dtype(dt)], n, m, p, Vs.ctypes.data, Ws.ctypes.data, n, c_pairs.ctypes.data, theta, out.ctypes.data))
    store = lambda a: a.astype(dt).astype(wide)  # noqa: E731 -- rounding to the storage type
    Vp, Wp = V0 @ Cw, W0 @ Cw
    r = store((W0 - theta * V0) @ Cw[:, 0])
    tol = 1e-12 if np.dtype(dt).itemsize >= 8 and np.dtype(dt) != np.complex64 else 2e-5
    scale = np.linalg.norm(W0)
    assert np.allclose(Ws[:, :p], Wp, rtol=0, atol=tol * scale)
    assert np.allclose(Vs[:, :p], Vp, rtol=0, atol=tol)
    sums = out[:2 * (1 + p)].reshape(1 + p, 2)
    assert abs(sums[0, 0] - np.vdot(r, r).real) <= tol * max(1.0, np.vdot(r, r).real)
    d = store(Vp).conj().T @ r
    got_d = sums[1:, 0] + 1j * sums[1:, 1]
    assert np.allclose(got_d, d, rtol=0, atol=tol * np.linalg.norm(r) + 1e-300)
    t = store(r - store(Vp) @ got_d) if cplx else store(r - store(Vp).real @ got_d.real)
    nt = np.vdot(t, t).real
    assert abs(out[2 * (1 + p)] - nt) <= 10 * tol * max(nt, 1e-300)
    assert abs(out[2 * (1 + p) + 1] - nt / np.vdot(r, r).real) <= 10 * tol
    assert out[2 * (1 + p) + 2] == (1.0 if nt / np.vdot(r, r).real < 0.5 else 0.0)
    assert np.allclose(Vs[:, p], t / np.sqrt(nt), rtol=0, atol=10 * tol)
    # the restarted basis stays orthonormal and the new direction is orthogonal to it
    G = Vs[:, :p + 1].astype(wide).conj().T @ Vs[:, :p + 1].astype(wide)
    assert np.allclose(G, np.eye(p + 1), atol=50 * tol)
