"""ctypes wrapper around ``oracle/_build/liboracle.so`` (see oracle.cpp for provenance).

TEST INFRASTRUCTURE ONLY -- never imported by the product package.
The Python surface mirrors the reference's FFI layer (src/SpinED/Internal.hs): a basis is made
from (number_spins, hamming_weight, spin_inversion, symmetries), built, queried for its
representatives; an operator is a list of (matrix, sites) terms bound to a basis and applied to
column-major blocks.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

DTYPES = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.complex64): 2, np.dtype(np.complex128): 3}


class OracleError(RuntimeError):
    def __init__(self, code: int):
        super().__init__(f"oracle error code {code}")
        self.code = code


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (g++ -fopenmp)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "oracle.cpp")
    ):
        env = dict(os.environ)
        env.pop("CXX", None)
        subprocess.check_call(["make", "-C", _HERE, "-s"], env=env)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        u64, vp, ci = C.c_uint64, C.c_void_p, C.c_int
        L.orc_basis_new.restype = vp
        L.orc_basis_new.argtypes = [ci, ci, ci, ci, vp, vp, C.POINTER(ci)]
        L.orc_basis_free.argtypes = [vp]
        L.orc_basis_use_naive.argtypes = [vp, ci]
        L.orc_basis_group_size.restype = u64
        L.orc_basis_group_size.argtypes = [vp]
        L.orc_basis_group_element.argtypes = [vp, u64, vp, C.POINTER(C.c_double)]
        L.orc_periodicity.restype = ci
        L.orc_periodicity.argtypes = [ci, vp]
        L.orc_basis_build.argtypes = [vp]
        L.orc_basis_build_unsafe.argtypes = [vp, u64, vp]
        L.orc_basis_adopt_lazy.argtypes = [vp, u64, vp]
        L.orc_basis_build_range.restype = u64
        L.orc_basis_build_range.argtypes = [vp, u64, u64, u64, vp, vp]
        L.orc_sector_candidates.restype = u64
        L.orc_sector_candidates.argtypes = [vp]
        L.orc_operator_matmat_list.argtypes = [vp, ci, u64, vp, u64, vp, vp, C.POINTER(u64)]
        L.orc_basis_size.restype = u64
        L.orc_basis_size.argtypes = [vp]
        L.orc_basis_states.argtypes = [vp, vp]
        L.orc_basis_norms.argtypes = [vp, vp]
        L.orc_state_info.argtypes = [vp, u64, C.POINTER(u64), vp, C.POINTER(C.c_double)]
        L.orc_basis_index.restype = C.c_longlong
        L.orc_basis_index.argtypes = [vp, u64]
        L.orc_apply_permutation.restype = u64
        L.orc_apply_permutation.argtypes = [ci, vp, u64, ci]
        L.orc_operator_new.restype = vp
        L.orc_operator_new.argtypes = [vp]
        L.orc_operator_free.argtypes = [vp]
        L.orc_operator_add_term.argtypes = [vp, ci, vp, ci, vp]
        L.orc_operator_is_real.argtypes = [vp]
        L.orc_operator_matmat.argtypes = [vp, ci, u64, u64, vp, u64, vp, u64, C.POINTER(u64)]
        L.orc_operator_matmat_rows.argtypes = [vp, ci, u64, u64, vp, u64, vp, u64, C.POINTER(u64), u64, u64, u64]
        L.orc_operator_expectation.argtypes = [vp, ci, u64, u64, vp, u64, vp]
        L.orc_num_threads.restype = ci
        L.orc_set_num_threads.argtypes = [ci]
        _lib = L
    return _lib


def _check(code: int):
    if code != 0:
        raise OracleError(code)


def periodicity(perm) -> int:
    p = np.ascontiguousarray(perm, dtype=np.int32)
    return int(lib().orc_periodicity(len(p), p.ctypes.data))


def apply_permutation(perm, x: int, naive: bool = False) -> int:
    p = np.ascontiguousarray(perm, dtype=np.int32)
    return int(lib().orc_apply_permutation(len(p), p.ctypes.data, x, int(naive)))


class Basis:
    def __init__(self, number_spins: int, hamming_weight=None, spin_inversion=None, symmetries=()):
        for s in symmetries:
            if len(s["permutation"]) != number_spins:
                raise OracleError(6)
        perms = np.ascontiguousarray([s["permutation"] for s in symmetries], dtype=np.int32).reshape(
            len(symmetries), number_spins
        )
        sectors = np.ascontiguousarray([s["sector"] for s in symmetries], dtype=np.int32)
        err = C.c_int(0)
        self._h = lib().orc_basis_new(
            number_spins,
            -1 if hamming_weight is None else hamming_weight,
            0 if spin_inversion is None else spin_inversion,
            len(symmetries),
            perms.ctypes.data if len(symmetries) else None,
            sectors.ctypes.data if len(symmetries) else None,
            C.byref(err),
        )
        if not self._h:
            raise OracleError(err.value)
        self.number_spins = number_spins
        self.hamming_weight = hamming_weight
        self.spin_inversion = spin_inversion

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_basis_free(self._h)
            self._h = None

    def use_naive(self, flag=True):
        lib().orc_basis_use_naive(self._h, int(flag))

    @property
    def group_size(self) -> int:
        return int(lib().orc_basis_group_size(self._h))

    def group_elements(self):
        out = []
        for i in range(self.group_size):
            p = np.zeros(self.number_spins, dtype=np.int32)
            ph = C.c_double(0)
            lib().orc_basis_group_element(self._h, i, p.ctypes.data, C.byref(ph))
            out.append((p.copy(), ph.value))
        return out

    def build(self, representatives=None):
        if representatives is None:
            _check(lib().orc_basis_build(self._h))
        else:
            r = np.ascontiguousarray(representatives, dtype=np.uint64)
            _check(lib().orc_basis_build_unsafe(self._h, len(r), r.ctypes.data))
        return self

    def adopt_lazy(self, representatives):
        """Adopt sorted representatives WITHOUT the O(N |G|) norm pass: norms are derived per use.
        For bounded samples of sectors with ~10^9 states (bench.py parity checks at size)."""
        r = np.ascontiguousarray(representatives, dtype=np.uint64)
        _check(lib().orc_basis_adopt_lazy(self._h, len(r), r.ctypes.data))
        return self

    @property
    def sector_candidates(self) -> int:
        """number of candidate words of the sector (C(n, hw) or 2^n) the enumeration walks"""
        return int(lib().orc_sector_candidates(self._h))

    def build_range(self, rank_lo: int, rank_hi: int):
        """Independent enumeration of the candidates of combinatorial rank [rank_lo, rank_hi):
        returns (representatives found, first candidate word, one-past-last candidate word or 2^64-1)."""
        cap = max(1, rank_hi - rank_lo)
        out = np.zeros(cap, dtype=np.uint64)
        rng = np.zeros(2, dtype=np.uint64)
        cnt = int(lib().orc_basis_build_range(self._h, rank_lo, rank_hi, cap, out.ctypes.data, rng.ctypes.data))
        return out[:cnt].copy(), int(rng[0]), int(rng[1])

    @property
    def number_states(self) -> int:
        return int(lib().orc_basis_size(self._h))

    @property
    def states(self) -> np.ndarray:
        out = np.zeros(self.number_states, dtype=np.uint64)
        lib().orc_basis_states(self._h, out.ctypes.data)
        return out

    @property
    def norms(self) -> np.ndarray:
        out = np.zeros(self.number_states, dtype=np.float64)
        lib().orc_basis_norms(self._h, out.ctypes.data)
        return out

    def state_info(self, x: int):
        rep = C.c_uint64(0)
        chi = np.zeros(2)
        norm = C.c_double(0)
        lib().orc_state_info(self._h, x, C.byref(rep), chi.ctypes.data, C.byref(norm))
        return int(rep.value), complex(chi[0], chi[1]), norm.value

    def index(self, rep: int) -> int:
        return int(lib().orc_basis_index(self._h, rep))


class Operator:
    def __init__(self, basis: Basis, terms):
        """terms: iterable of dicts {matrix: 2^k x 2^k (complex), sites: [[..k..], ...]}"""
        self.basis = basis
        self._h = lib().orc_operator_new(basis._h)
        for t in terms:
            m = np.ascontiguousarray(t["matrix"], dtype=np.complex128)
            sites = np.ascontiguousarray(t["sites"], dtype=np.int32)
            k = sites.shape[1]
            if m.shape != (1 << k, 1 << k):
                raise OracleError(2)
            _check(lib().orc_operator_add_term(self._h, k, m.ctypes.data, sites.shape[0], sites.ctypes.data))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_operator_free(self._h)
            self._h = None

    @property
    def is_real(self) -> bool:
        return bool(lib().orc_operator_is_real(self._h))

    def matmat(self, x: np.ndarray, count: bool = False):
        """x: (N,) or column-major (N, block) -> y of the same shape/dtype."""
        x = np.asarray(x)
        squeeze = x.ndim == 1
        X = np.asfortranarray(x.reshape(len(x), -1))
        Y = np.zeros_like(X, order="F")
        n_off = C.c_uint64(0)
        _check(
            lib().orc_operator_matmat(
                self._h, DTYPES[X.dtype], X.shape[0], X.shape[1], X.ctypes.data, X.shape[0], Y.ctypes.data, Y.shape[0],
                C.byref(n_off),
            )
        )
        Y = Y[:, 0] if squeeze else Y
        return (Y, int(n_off.value)) if count else Y

    def matmat_rows(self, x: np.ndarray, y: np.ndarray, row_lo: int, row_hi: int, row_stride: int = 1) -> int:
        """Rows row_lo, row_lo+stride, ... < row_hi of y = H x (1-D x, y); returns the number of
        off-diagonal elements visited.  Bounded-sample entry for CPU timing."""
        n_off = C.c_uint64(0)
        _check(lib().orc_operator_matmat_rows(self._h, DTYPES[x.dtype], len(x), 1, x.ctypes.data, len(x), y.ctypes.data,
                                              len(y), C.byref(n_off), row_lo, row_hi, row_stride))
        return int(n_off.value)

    def matmat_list(self, x: np.ndarray, rows: np.ndarray):
        """(H x)[rows] for an explicit list of rows (1-D x); returns (values, off-diagonal elements visited)."""
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        out = np.zeros(len(rows), dtype=x.dtype)
        n_off = C.c_uint64(0)
        _check(lib().orc_operator_matmat_list(self._h, DTYPES[x.dtype], len(x), x.ctypes.data, len(rows), rows.ctypes.data,
                                              out.ctypes.data, C.byref(n_off)))
        return out, int(n_off.value)

    def count_offdiag(self) -> int:
        """E = number of off-diagonal term applications with non-zero target norm."""
        n = self.basis.number_states
        x = np.zeros((n, 1), dtype=np.complex128 if not self.is_real else np.float64, order="F")
        n_off = C.c_uint64(0)
        _check(lib().orc_operator_matmat(self._h, DTYPES[x.dtype], n, 1, x.ctypes.data, n, None, n, C.byref(n_off)))
        return int(n_off.value)

    def expectation(self, x: np.ndarray) -> np.ndarray:
        X = np.asfortranarray(np.asarray(x).reshape(len(x), -1))
        out = np.zeros(X.shape[1], dtype=np.complex128)
        _check(lib().orc_operator_expectation(self._h, DTYPES[X.dtype], X.shape[0], X.shape[1], X.ctypes.data, X.shape[0], out.ctypes.data))
        return out

    def to_dense(self) -> np.ndarray:
        n = self.basis.number_states
        dt = np.float64 if self.is_real else np.complex128
        return self.matmat(np.eye(n, dtype=dt, order="F"))


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int):
    lib().orc_set_num_threads(n)
