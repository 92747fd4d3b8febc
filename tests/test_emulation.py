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
                       "chain_10_inv_nohw", "chain_8_chiral_3site", "ring_4site_nosym", "chain_40_hw3_k"]


def _emulate(op, reps, stab, world, rank, x, ncols=1):
    n = len(reps)
    rd = ffi.rowDistribution(n, world, rank)
    n_local = int(rd.n_local)
    dt = x.dtype
    outs = [np.full(max(n_local, 1), 7.0, dtype=dt) for _ in range(3)]
    block = np.full((max(n_local, 1), max(ncols, 1)), 7.0, dtype=dt, order="F")
    stats = (C.c_uint64 * 5)()
    ffi.checkStatus(ffi.emulLib().sped_selftest_emulate_matvec(
        op._ptr, n, reps.ctypes.data, stab.ctypes.data, world, rank, ffi.DTYPE_TAGS[np.dtype(dt)], x.ctypes.data,
        outs[0].ctypes.data, outs[1].ctypes.data, outs[2].ctypes.data, stats, ncols, block.ctypes.data))
    rows = rd.local_rows().astype(np.int64)
    if ncols > 1:
        return rows, [o[:n_local] for o in outs], [int(v) for v in stats], block[:n_local]
    return rows, [o[:n_local] for o in outs], [int(v) for v in stats]


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
