"""YAML input schema and object construction -- the Python mirror of the config half of
/root/reference/src/SpinED.hs: ``SymmetrySpec`` (:81-92), ``BasisSpec`` (:94-112), complex numbers
as ``[re, im]`` (:114-122), ``InteractionSpec`` (:124-131), ``OperatorSpec`` (:133-140),
``Datatype`` (:142-156), ``ConfigSpec`` and its defaults (:158-173), ``toBasis`` / ``toOperator`` /
``toConfig`` (:178-248), ``readConfig`` (:250-258).
"""
from __future__ import annotations

import dataclasses
from typing import Any, List, Optional

from . import ffi
from .ffi import SpinEDException


@dataclasses.dataclass
class SymmetrySpec:
    permutation: List[int]
    sector: int


@dataclasses.dataclass
class BasisSpec:
    number_spins: int
    hamming_weight: Optional[int]
    spin_inversion: Optional[int]
    symmetries: List[SymmetrySpec]


@dataclasses.dataclass
class InteractionSpec:
    matrix: List[List[Any]]
    sites: List[List[int]]


@dataclasses.dataclass
class OperatorSpec:
    name: str
    terms: List[InteractionSpec]


@dataclasses.dataclass
class ConfigSpec:
    basis: BasisSpec
    hamiltonian: OperatorSpec
    observables: List[OperatorSpec]
    output: str = "exact_diagonalization_result.h5"
    number_vectors: int = 1
    precision: float = 0.0
    datatype: str = "float64"
    # PRIMME's own defaults are represented by 0 ("let the solver choose"), like primmeDefaults
    max_primme_basis_size: int = 0
    max_primme_block_size: int = 0
    min_primme_restart_size: int = 0


def _require(d: dict, key: str, what: str):
    if not isinstance(d, dict) or key not in d:
        raise SpinEDException(f"parsing {what} failed: key {key!r} not found")
    return d[key]


def parseSymmetry(d) -> SymmetrySpec:
    return SymmetrySpec([int(v) for v in _require(d, "permutation", "symmetry")], int(_require(d, "sector", "symmetry")))


def parseBasis(d) -> BasisSpec:
    return BasisSpec(
        int(_require(d, "number_spins", "basis")),
        None if d.get("hamming_weight") is None else int(d["hamming_weight"]),
        None if d.get("spin_inversion") is None else int(d["spin_inversion"]),
        [parseSymmetry(s) for s in _require(d, "symmetries", "basis")],
    )


def parseInteraction(d) -> InteractionSpec:
    return InteractionSpec(_require(d, "matrix", "interaction"), _require(d, "sites", "interaction"))


def parseOperator(d) -> OperatorSpec:
    return OperatorSpec(str(_require(d, "name", "operator")), [parseInteraction(t) for t in _require(d, "terms", "operator")])


def parseDatatype(v: str) -> str:
    if str(v).lower() not in ("float32", "float64"):
        raise SpinEDException(f"parsing Datatype failed, expected either 'float32' or 'float64', but got '{v}'")
    return str(v).lower()


def parseConfig(d: dict) -> ConfigSpec:
    spec = ConfigSpec(
        parseBasis(_require(d, "basis", "config")),
        parseOperator(_require(d, "hamiltonian", "config")),
        [parseOperator(o) for o in _require(d, "observables", "config")],
    )
    for key in ("output", "number_vectors", "precision", "max_primme_basis_size", "max_primme_block_size", "min_primme_restart_size"):
        if d.get(key) is not None:
            setattr(spec, key, type(getattr(spec, key))(d[key]))
    if d.get("datatype") is not None:
        spec.datatype = parseDatatype(d["datatype"])
    return spec


def readConfig(path: str) -> ConfigSpec:
    import yaml

    with open(path) as f:
        return parseConfig(yaml.safe_load(f))


def toSymmetry(s: SymmetrySpec) -> ffi.Symmetry:
    return ffi.mkSymmetry(s.permutation, s.sector)


def toBasis(spec: BasisSpec, log=None) -> ffi.SpinBasis:
    if spec.number_spins > 64:
        raise SpinEDException(
            f"invalid number_spins: {spec.number_spins}; exact diagonalization is not feasible for systems larger than 64 spins"
        )
    group = ffi.mkGroup([toSymmetry(s) for s in spec.symmetries])
    if log:
        log(f"Symmetry group contains {ffi.getGroupSize(group)} elements")
    basis = ffi.mkBasis(group, spec.number_spins, spec.hamming_weight, spec.spin_inversion)
    basis.group = group
    return basis


def toInteraction(spec: InteractionSpec) -> ffi.Interaction:
    return ffi.mkInteraction(spec.matrix, spec.sites)


@dataclasses.dataclass
class Operator:
    operatorName: str
    operatorObject: ffi.Operator


def toOperator(basis: ffi.SpinBasis, spec: OperatorSpec) -> Operator:
    return Operator(spec.name, ffi.mkOperator(basis, [toInteraction(t) for t in spec.terms]))


@dataclasses.dataclass
class UserConfig:
    cBasis: ffi.SpinBasis
    cHamiltonian: Operator
    cObservables: List[Operator]
    cOutput: str
    cNumEvals: int
    cEps: float
    cDatatype: str
    cMaxBasisSize: int
    cMaxBlockSize: int
    cMinRestartSize: int


def toConfig(spec: ConfigSpec, log=None) -> UserConfig:
    basis = toBasis(spec.basis, log)
    hamiltonian = toOperator(basis, spec.hamiltonian)
    observables = [toOperator(basis, o) for o in spec.observables]
    return UserConfig(basis, hamiltonian, observables, spec.output, spec.number_vectors, spec.precision, spec.datatype,
                      spec.max_primme_basis_size, spec.max_primme_block_size, spec.min_primme_restart_size)
