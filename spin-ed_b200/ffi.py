"""ctypes binding of libsped.so -- the Python mirror of the reference's FFI layer
(/root/reference/src/SpinED/Internal.hs).  Function and class names follow that module:
``mkSymmetry``/``Symmetry`` (:99-116), ``mkGroup``/``SymmetryGroup`` (:141-150), ``mkBasis``/
``SpinBasis`` (:199-228), ``buildBasis`` (:230-236), ``basisGetStates`` (:238-244),
``getNumberStates`` (:247-254), ``mkInteraction'``/``Interaction`` (:293-362), ``mkOperator``/
``Operator'`` (:391-402), ``inplaceApply`` (:411-429), ``apply`` (:431-435), ``expectation``
(:437-452).  Error behaviour is the same: a non-zero status raises ``LatticeSymmetriesException``
(code, message) (:48-65); argument errors caught before the FFI raise ``SpinEDException``.

There is no CPU fallback: every compute call goes to the CUDA library and fails loudly if the
library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsped.so")
_lib = None

F32, F64, C64, C128 = 0, 1, 2, 3
DTYPE_TAGS = {np.dtype(np.float32): F32, np.dtype(np.float64): F64, np.dtype(np.complex64): C64, np.dtype(np.complex128): C128}
TAG_DTYPES = {v: k for k, v in DTYPE_TAGS.items()}


class LatticeSymmetriesException(RuntimeError):
    """Mirror of ``LatticeSymmetriesException {eCode, eMessage}`` (Internal.hs:28-31)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.eCode = code
        self.eMessage = message


class SpinEDException(RuntimeError):
    """Mirror of ``SpinEDException`` (Internal.hs:33-36)."""


class sped_eigh_info(C.Structure):
    _fields_ = [
        ("iteration", C.c_int), ("basis_size", C.c_int), ("number_converged", C.c_int), ("number_evals", C.c_int),
        ("number_matvecs", C.c_uint64), ("evals", C.POINTER(C.c_double)), ("rnorms", C.POINTER(C.c_double)),
        ("elapsed_seconds", C.c_double),
    ]


class sped_eigh_stats(C.Structure):
    _fields_ = [
        ("matvecs", C.c_uint64), ("iterations", C.c_int), ("restarts", C.c_int), ("seconds_total", C.c_double),
        ("seconds_matvec", C.c_double), ("seconds_ortho", C.c_double), ("seconds_residual", C.c_double),
        ("seconds_restart", C.c_double), ("seconds_project", C.c_double),
    ]


MONITOR_FN = C.CFUNCTYPE(C.c_int, C.POINTER(sped_eigh_info), C.c_void_p)

# name -> (restype, argtypes); the list is the contract checked against include/sped.h by the tests
_vp, _u64, _ci, _cu, _pp = C.c_void_p, C.c_uint64, C.c_int, C.c_uint, C.POINTER(C.c_void_p)
SIGNATURES = {
    "ls_error_to_string": (C.c_void_p, [_ci]),
    "ls_destroy_string": (None, [_vp]),
    "ls_enable_logging": (None, []),
    "ls_disable_logging": (None, []),
    "ls_create_symmetry": (_ci, [_pp, _cu, _vp, _cu]),
    "ls_destroy_symmetry": (None, [_vp]),
    "ls_get_sector": (_cu, [_vp]),
    "ls_get_phase": (C.c_double, [_vp]),
    "ls_get_periodicity": (_cu, [_vp]),
    "ls_create_group": (_ci, [_pp, _cu, _vp]),
    "ls_destroy_group": (None, [_vp]),
    "ls_get_group_size": (_cu, [_vp]),
    "ls_create_spin_basis": (_ci, [_pp, _vp, _cu, _ci, _ci]),
    "ls_destroy_spin_basis": (None, [_vp]),
    "ls_build": (_ci, [_vp]),
    "ls_build_unsafe": (_ci, [_vp, _u64, _vp]),
    "ls_get_number_states": (_ci, [_vp, C.POINTER(_u64)]),
    "ls_get_states": (_ci, [_pp, _vp]),
    "ls_states_get_data": (_vp, [_vp]),
    "ls_states_get_size": (_u64, [_vp]),
    "ls_destroy_states": (None, [_vp]),
    "ls_create_interaction1": (_ci, [_pp, _vp, _cu, _vp]),
    "ls_create_interaction2": (_ci, [_pp, _vp, _cu, _vp]),
    "ls_create_interaction3": (_ci, [_pp, _vp, _cu, _vp]),
    "ls_create_interaction4": (_ci, [_pp, _vp, _cu, _vp]),
    "ls_interaction_is_real": (C.c_bool, [_vp]),
    "ls_destroy_interaction": (None, [_vp]),
    "ls_create_operator": (_ci, [_pp, _vp, _cu, _vp]),
    "ls_destroy_operator": (None, [_vp]),
    "ls_operator_is_real": (C.c_bool, [_vp]),
    "ls_operator_matmat": (_ci, [_vp, _ci, _u64, _u64, _vp, _u64, _vp, _u64]),
    "ls_operator_expectation": (_ci, [_vp, _ci, _u64, _u64, _vp, _u64, _vp]),
    "sped_version": (C.c_char_p, []),
    "sped_device_count": (_ci, [C.POINTER(_ci)]),
    "sped_set_device": (_ci, [_ci]),
    "sped_kernel_launches": (_u64, []),
    "sped_comm_unique_id": (_ci, [_vp]),
    "sped_comm_init": (_ci, [_ci, _ci, _vp]),
    "sped_comm_finalize": (_ci, []),
    "sped_comm_rank": (_ci, []),
    "sped_comm_size": (_ci, []),
    "sped_row_distribution": (None, [_u64, _ci, _ci, _vp]),
    "sped_dist_local_to_global": (_u64, [_vp, _u64]),
    "sped_dist_global_to_position": (_u64, [_vp, _u64]),
    "sped_basis_build_seconds": (_ci, [_vp, C.POINTER(C.c_double)]),
    "sped_basis_row_distribution": (_ci, [_vp, _vp]),
    "sped_basis_device_states": (_ci, [_vp, _pp]),
    "sped_basis_norms": (_ci, [_vp, _vp]),
    "sped_basis_state_info": (_ci, [_vp, _u64, _vp, _vp, _vp, _vp]),
    "sped_basis_program_stats": (_ci, [_vp, C.POINTER(_cu), C.POINTER(_cu), C.POINTER(_cu)]),
    "sped_operator_matmat_device": (_ci, [_vp, _ci, _u64, _vp, _u64, _vp, _u64, _vp]),
    "sped_operator_matvec_sharded": (_ci, [_vp, _ci, _vp, _vp, _vp, _vp]),
    "sped_operator_matmat_local": (_ci, [_vp, _ci, _u64, _vp, _u64, _vp, _u64]),
    "sped_operator_count_elements": (_ci, [_vp, C.POINTER(_u64), C.POINTER(_u64)]),
    "sped_operator_diagonal": (_ci, [_vp, _vp]),
    "sped_operator_set_cache": (_ci, [_vp, _ci]),
    "sped_operator_cache_info": (_ci, [_vp, C.POINTER(_ci), C.POINTER(_u64), C.POINTER(C.c_double)]),
    "sped_eigh": (_ci, [_vp, _ci, _u64, C.c_double, _ci, _ci, _ci, _vp, _vp, _vp, MONITOR_FN, _vp]),
    "sped_operator_release_workspace": (_ci, [_vp]),
    "sped_eigh_last_stats": (_ci, [_vp, C.POINTER(sped_eigh_stats)]),
    "sped_selftest_small_eigh": (_ci, [_ci, _vp, _vp, _vp]),
    "sped_selftest_program": (_ci, [_vp, _u64, _vp, _vp, _vp, _vp]),
    "sped_selftest_burnside": (_ci, [_vp, C.POINTER(_u64)]),
    "sped_selftest_jit_source": (_ci, [_vp, _vp, _u64, C.POINTER(_u64)]),
    "sped_selftest_jit_compile": (_ci, [_vp, _ci, _ci, C.POINTER(_u64)]),
}
# test-only library (host emulation of the kernels, csrc/emul.cpp): not part of the product
EMUL_LIB_PATH = os.path.join(_HERE, "lib", "libsped_emul.so")
EMUL_SIGNATURES = {
    "sped_selftest_emulate_matvec": (_ci, [_vp, _u64, _vp, _vp, _ci, _ci, _ci, _vp, _vp, _vp, _vp, _vp, C.c_uint, _vp]),
    "sped_selftest_emulate_restart": (_ci, [_ci, _u64, _ci, _ci, _vp, _vp, _u64, _vp, C.c_double, _vp]),
}
_emul_lib = None


def emulLib():
    """libsped_emul.so: the kernels' device sources compiled by the host compiler (tests only)."""
    global _emul_lib
    if _emul_lib is None:
        lib()
        L = C.CDLL(EMUL_LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in EMUL_SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _emul_lib = L
    return _emul_lib


def lib():
    """Load libsped.so (built in-tree by ``__graft_entry__.build()``); no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()')"
            )
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def getErrorMessage(code: int) -> str:
    p = lib().ls_error_to_string(code)
    try:
        return C.cast(p, C.c_char_p).value.decode()
    finally:
        lib().ls_destroy_string(p)


def checkStatus(code: int):
    if code != 0:
        raise LatticeSymmetriesException(code, getErrorMessage(code))


def _mkObject(f):
    out = C.c_void_p()
    checkStatus(f(C.byref(out)))
    return out


class _Handle:
    _destroy = None

    def __init__(self, ptr):
        self._ptr = ptr

    def __del__(self):
        p, self._ptr = getattr(self, "_ptr", None), None
        if p and _lib is not None:
            getattr(_lib, self._destroy)(p)


class Symmetry(_Handle):
    _destroy = "ls_destroy_symmetry"


def mkSymmetry(permutation, sector: int) -> Symmetry:
    if any(int(p) < 0 for p in permutation):
        raise SpinEDException(f"invalid permutation: {list(permutation)}; indices must be non-negative")
    if sector < 0:
        raise SpinEDException(f"invalid sector: {sector}; expected a non-negative number")
    p = np.ascontiguousarray(permutation, dtype=np.uint32)
    return Symmetry(_mkObject(lambda out: lib().ls_create_symmetry(out, len(p), p.ctypes.data, sector)))


def getSector(s: Symmetry) -> int:
    return int(lib().ls_get_sector(s._ptr))


def getPeriodicity(s: Symmetry) -> int:
    return int(lib().ls_get_periodicity(s._ptr))


def getPhase(s: Symmetry) -> float:
    return float(lib().ls_get_phase(s._ptr))


class SymmetryGroup(_Handle):
    _destroy = "ls_destroy_group"


def mkGroup(symmetries) -> SymmetryGroup:
    arr = (C.c_void_p * max(1, len(symmetries)))(*[s._ptr for s in symmetries])
    return SymmetryGroup(_mkObject(lambda out: lib().ls_create_group(out, len(symmetries), arr)))


def getGroupSize(g: SymmetryGroup) -> int:
    return int(lib().ls_get_group_size(g._ptr))


class SpinBasis(_Handle):
    _destroy = "ls_destroy_spin_basis"


def mkBasis(group: SymmetryGroup, numberSpins: int, hammingWeight=None, spinInversion=None) -> SpinBasis:
    if numberSpins <= 0:
        raise SpinEDException(f"invalid number of spins: {numberSpins}; expected a positive number")
    if hammingWeight is not None and hammingWeight < 0:
        raise SpinEDException(f"invalid Hamming weight: {hammingWeight}; expected a non-negative number")
    if spinInversion is not None and spinInversion not in (1, -1):
        raise SpinEDException(f"invalid value for spin inversion: {spinInversion}; expected either -1 or +1")
    hw = -1 if hammingWeight is None else hammingWeight
    inv = 0 if spinInversion is None else spinInversion
    b = SpinBasis(_mkObject(lambda out: lib().ls_create_spin_basis(out, group._ptr, numberSpins, hw, inv)))
    b.number_spins = numberSpins
    return b


def buildBasis(basis: SpinBasis, representatives=None):
    if representatives is None:
        checkStatus(lib().ls_build(basis._ptr))
    else:
        r = np.ascontiguousarray(representatives, dtype=np.uint64)
        checkStatus(lib().ls_build_unsafe(basis._ptr, len(r), r.ctypes.data))


def getNumberStates(basis: SpinBasis) -> int:
    out = C.c_uint64(0)
    checkStatus(lib().ls_get_number_states(basis._ptr, C.byref(out)))
    return int(out.value)


class _States(_Handle):
    _destroy = "ls_destroy_states"


def basisGetStates(basis: SpinBasis) -> np.ndarray:
    """Zero-copy view of the representatives; the ``ls_states`` object lives as long as the array."""
    st = _States(_mkObject(lambda out: lib().ls_get_states(out, basis._ptr)))
    n = int(lib().ls_states_get_size(st._ptr))
    if n == 0:
        return np.zeros(0, dtype=np.uint64)
    buf = (C.c_uint64 * n).from_address(lib().ls_states_get_data(st._ptr))
    arr = np.frombuffer(buf, dtype=np.uint64).view(_StatesArray)
    arr._owner = st  # the ls_states object is destroyed when the last view of the array goes away
    arr.flags.writeable = False
    return arr


class _StatesArray(np.ndarray):
    _owner = None


class Interaction(_Handle):
    _destroy = "ls_destroy_interaction"


def toMatrix(dim: int, rows) -> np.ndarray:
    rows = [list(r) for r in rows]
    if len(rows) != dim or any(len(r) != dim for r in rows):
        raise SpinEDException(f"invalid matrix: {rows}; expected a square matrix of dimension {dim}")
    out = np.zeros((dim, dim), dtype=np.complex128)
    for i, r in enumerate(rows):
        for j, v in enumerate(r):
            out[i, j] = complex(*v) if isinstance(v, (list, tuple)) else complex(v)
    return out


def mkInteraction(matrix, sites) -> Interaction:
    """``mkInteraction'`` for site tuples of uniform length 1..4 (Internal.hs:293-362)."""
    sites = [list(s) if isinstance(s, (list, tuple)) else [s] for s in sites]
    if not sites:
        raise SpinEDException("zero-point interactions (i.e. constant factors) are not supported")
    n = len(sites[0])
    if n not in (1, 2, 3, 4):
        raise SpinEDException(f"currently only 1-, 2-, 3-, and 4-point interactions are supported, but received n={n}")
    if any(len(s) != n for s in sites):
        raise SpinEDException(f"invalid sites: {sites}; expected an array of length-{n} tuples")
    m = np.ascontiguousarray(toMatrix(1 << n, matrix))
    flat = [int(v) for s in sites for v in s]
    if any(v < 0 for v in flat):
        raise SpinEDException("site indices must be all non-negative numbers")
    s16 = np.ascontiguousarray(flat, dtype=np.uint16)
    fn = getattr(lib(), f"ls_create_interaction{n}")
    return Interaction(_mkObject(lambda out: fn(out, m.ctypes.data, len(sites), s16.ctypes.data)))


def isRealInteraction(t: Interaction) -> bool:
    return bool(lib().ls_interaction_is_real(t._ptr))


class Operator(_Handle):
    """``Operator'`` (Internal.hs:369)."""

    _destroy = "ls_destroy_operator"


def mkOperator(basis: SpinBasis, terms) -> Operator:
    arr = (C.c_void_p * max(1, len(terms)))(*[t._ptr for t in terms])
    op = Operator(_mkObject(lambda out: lib().ls_create_operator(out, basis._ptr, len(terms), arr)))
    op.basis = basis
    return op


def isOperatorReal(op: Operator) -> bool:
    return bool(lib().ls_operator_is_real(op._ptr))


def _block(x: np.ndarray):
    """Column-major (size, blockSize) view + stride, like primme-hs's ``Block``."""
    x = np.asarray(x)
    if x.dtype not in DTYPE_TAGS:
        raise SpinEDException(f"unsupported datatype {x.dtype}")
    X = x.reshape(len(x), -1)
    if not X.flags.f_contiguous:
        X = np.asfortranarray(X)
    return X, X.shape[0], X.shape[1], max(X.strides[1] // X.itemsize, X.shape[0]) if X.shape[1] > 1 else X.shape[0]


def inplaceApply(op: Operator, x: np.ndarray, y: np.ndarray):
    X, size, block, xs = _block(x)
    if y.shape != x.shape or y.dtype != x.dtype:
        raise SpinEDException(f"dimensions of x and y do not match: {x.shape} != {y.shape}")
    if not (y.flags.f_contiguous or y.ndim == 1):
        raise SpinEDException("y must be column-major")
    checkStatus(lib().ls_operator_matmat(op._ptr, DTYPE_TAGS[X.dtype], size, block, X.ctypes.data, xs, y.ctypes.data, size))


def apply(op: Operator, x: np.ndarray) -> np.ndarray:
    x = np.asarray(x)
    y = np.zeros(x.shape, dtype=x.dtype, order="F")
    inplaceApply(op, x, y)
    return y


def expectation(op: Operator, x: np.ndarray) -> np.ndarray:
    X, size, block, xs = _block(x)
    out = np.zeros(block, dtype=np.complex128)
    checkStatus(lib().ls_operator_expectation(op._ptr, DTYPE_TAGS[X.dtype], size, block, X.ctypes.data, xs, out.ctypes.data))
    return out


# ---------------------------------------------------------------------------------------------
# group B: device-resident solver and plumbing
# ---------------------------------------------------------------------------------------------
def deviceCount() -> int:
    out = C.c_int(0)
    checkStatus(lib().sped_device_count(C.byref(out)))
    return out.value


def setDevice(i: int):
    checkStatus(lib().sped_set_device(i))


def kernelLaunches() -> int:
    return int(lib().sped_kernel_launches())


class RowDist(C.Structure):
    """``sped_row_dist``: block-cyclic row distribution (include/sped.h)."""

    _fields_ = [("n", C.c_uint64), ("n_local", C.c_uint64), ("chunk", C.c_uint64), ("world", C.c_uint32),
                ("rank", C.c_uint32), ("log2_block", C.c_uint32), ("reserved", C.c_uint32)]

    def local_to_global(self, local) -> np.ndarray:
        """Global row indices of local indices (vectorised mirror of sped_dist_local_to_global)."""
        i = np.asarray(local, dtype=np.uint64)
        if self.world == 1:
            return i
        lb = np.uint64(self.log2_block)
        mask = np.uint64((1 << self.log2_block) - 1)
        return (((i >> lb) * np.uint64(self.world) + np.uint64(self.rank)) << lb) + (i & mask)

    def global_to_position(self, rows) -> np.ndarray:
        g = np.asarray(rows, dtype=np.uint64)
        if self.world == 1:
            return g
        lb = np.uint64(self.log2_block)
        mask = np.uint64((1 << self.log2_block) - 1)
        blk = g >> lb
        return (blk % np.uint64(self.world)) * np.uint64(self.chunk) + ((blk // np.uint64(self.world)) << lb) + (g & mask)

    def local_rows(self) -> np.ndarray:
        return self.local_to_global(np.arange(self.n_local, dtype=np.uint64))


def rowDistribution(n: int, world: int, rank: int) -> RowDist:
    d = RowDist()
    lib().sped_row_distribution(n, world, rank, C.byref(d))
    return d


def commUniqueId() -> bytes:
    buf = C.create_string_buffer(128)
    checkStatus(lib().sped_comm_unique_id(buf))
    return buf.raw


def commInit(world: int, rank: int, unique_id: bytes):
    checkStatus(lib().sped_comm_init(world, rank, C.create_string_buffer(unique_id, 128)))


def commFinalize():
    checkStatus(lib().sped_comm_finalize())


def basisBuildSeconds(basis: SpinBasis) -> float:
    out = C.c_double(0)
    checkStatus(lib().sped_basis_build_seconds(basis._ptr, C.byref(out)))
    return out.value


def basisRowDistribution(basis: SpinBasis) -> RowDist:
    d = RowDist()
    checkStatus(lib().sped_basis_row_distribution(basis._ptr, C.byref(d)))
    return d


def basisNorms(basis: SpinBasis) -> np.ndarray:
    out = np.zeros(getNumberStates(basis), dtype=np.float64)
    checkStatus(lib().sped_basis_norms(basis._ptr, out.ctypes.data))
    return out


def basisStateInfo(basis: SpinBasis, states):
    s = np.ascontiguousarray(states, dtype=np.uint64)
    reps = np.zeros(len(s), dtype=np.uint64)
    chars = np.zeros(len(s), dtype=np.complex128)
    norms = np.zeros(len(s), dtype=np.float64)
    checkStatus(lib().sped_basis_state_info(basis._ptr, len(s), s.ctypes.data, reps.ctypes.data, chars.ctypes.data, norms.ctypes.data))
    return reps, chars, norms


def basisProgramStats(basis: SpinBasis):
    a, b, c = C.c_uint(0), C.c_uint(0), C.c_uint(0)
    checkStatus(lib().sped_basis_program_stats(basis._ptr, C.byref(a), C.byref(b), C.byref(c)))
    return {"steps": a.value, "rotate_mask_ops": b.value, "delta_swap_ops": c.value}


def operatorMatmatDevice(op: Operator, dtype_tag: int, block: int, x_ptr: int, x_stride: int, y_ptr: int, y_stride: int, stream: int = 0):
    checkStatus(lib().sped_operator_matmat_device(op._ptr, dtype_tag, block, x_ptr, x_stride, y_ptr, y_stride, stream))


def operatorMatvecSharded(op: Operator, dtype_tag: int, x_local_ptr: int, y_local_ptr: int, x_replicated_ptr: int, stream: int = 0):
    """One column from this rank's shard: all-gather (library stream) overlapped with the local-source pass."""
    checkStatus(lib().sped_operator_matvec_sharded(op._ptr, dtype_tag, x_local_ptr, y_local_ptr, x_replicated_ptr, stream))


def applyLocal(op: Operator, x_local: np.ndarray, y_local: np.ndarray):
    """Row-sharded host-pointer product: this rank's rows of x in, this rank's rows of y out
    (1-D arrays or column-major blocks of n_local rows)."""
    xb, yb = np.asarray(x_local), np.asarray(y_local)
    if xb.dtype != yb.dtype or xb.shape != yb.shape:
        raise SpinEDException("x and y must have the same shape and dtype")
    rows = xb.shape[0] if xb.ndim else 0
    cols = xb.shape[1] if xb.ndim == 2 else 1
    if xb.ndim == 2 and not (xb.flags.f_contiguous and yb.flags.f_contiguous):
        raise SpinEDException("blocks must be column-major")
    checkStatus(lib().sped_operator_matmat_local(op._ptr, DTYPE_TAGS[xb.dtype], cols, xb.ctypes.data, max(rows, 1),
                                                 yb.ctypes.data, max(rows, 1)))


def operatorCountElements(op: Operator):
    r, e = C.c_uint64(0), C.c_uint64(0)
    checkStatus(lib().sped_operator_count_elements(op._ptr, C.byref(r), C.byref(e)))
    return int(r.value), int(e.value)


def operatorSetCache(op: Operator, mode: int):
    """-1: auto (default), 0: always matrix-free, 1: cache whenever it fits."""
    checkStatus(lib().sped_operator_set_cache(op._ptr, mode))


def operatorCacheInfo(op: Operator) -> dict:
    r, b, s = C.c_int(0), C.c_uint64(0), C.c_double(0)
    checkStatus(lib().sped_operator_cache_info(op._ptr, C.byref(r), C.byref(b), C.byref(s)))
    return {"ready": bool(r.value), "bytes": int(b.value), "build_seconds": s.value}


def operatorDiagonal(op: Operator) -> np.ndarray:
    out = np.zeros(basisRowDistribution(op.basis).n_local, dtype=np.float64)
    checkStatus(lib().sped_operator_diagonal(op._ptr, out.ctypes.data))
    return out


def eigh(op: Operator, dtype, numEvals=1, eps=0.0, maxBasisSize=0, maxBlockSize=0, minRestartSize=0, monitor=None,
         want_vectors=True):
    """Replacement of ``Numeric.PRIMME.eigh`` (SpinED.hs:404): -> (evals, evecs (N, k) column-major, rnorms)."""
    dtype = np.dtype(dtype)
    tag = DTYPE_TAGS[dtype]
    n = getNumberStates(op.basis)
    evals = np.zeros(numEvals, dtype=np.float64)
    rnorms = np.zeros(numEvals, dtype=np.float64)
    evecs = np.zeros((n, numEvals), dtype=dtype, order="F") if want_vectors else None

    def _cb(info_p, _ctx):
        info = info_p.contents
        k = info.number_evals
        return int(bool(monitor({
            "iteration": info.iteration, "basis_size": info.basis_size, "number_converged": info.number_converged,
            "number_matvecs": info.number_matvecs, "evals": [info.evals[i] for i in range(k)],
            "rnorms": [info.rnorms[i] for i in range(k)], "elapsed_seconds": info.elapsed_seconds,
        })))

    cb = MONITOR_FN(_cb) if monitor is not None else C.cast(None, MONITOR_FN)
    rc = lib().sped_eigh(op._ptr, tag, numEvals, eps, maxBasisSize, maxBlockSize, minRestartSize, evals.ctypes.data,
                         evecs.ctypes.data if want_vectors else None, rnorms.ctypes.data, cb, None)
    checkStatus(rc)
    return evals, evecs, rnorms


def operatorReleaseWorkspace(op: Operator):
    """Free the device workspace sped_eigh keeps on the operator between calls."""
    checkStatus(lib().sped_operator_release_workspace(op._ptr))


def eighLastStats(op: Operator) -> dict:
    st = sped_eigh_stats()
    checkStatus(lib().sped_eigh_last_stats(op._ptr, C.byref(st)))
    return {f: getattr(st, f) for f, _ in st._fields_}
