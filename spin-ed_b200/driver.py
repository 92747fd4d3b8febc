"""The run sequence of the reference's executable (/root/reference/app/Main.hs:14-27) on top of
libsped.so -- a Python stand-in for the Haskell driver, which cannot be compiled here:

    prepareOutputFile      SpinED.hs:289-294   groups /basis /hamiltonian /observables /_workspace
    buildRepresentatives   SpinED.hs:315-337   resume from /basis/representatives if present
    isOperatorReal + withDatatype  SpinED.hs:339-352
    diagonalize            SpinED.hs:370-411   -> /hamiltonian/{eigenvalues,eigenvectors,residuals}
    computeExpectations    SpinED.hs:296-313   -> /observables/<name>

    python -m spin_ed_b200.driver [--debug] input.yaml
"""
from __future__ import annotations

import argparse
import sys
import time

import numpy as np

from . import config as C
from . import ffi, hdf5

observablesPath, basisPath, hamiltonianPath, workspacePath = "/observables", "/basis", "/hamiltonian", "/_workspace"


def logInfo(msg: str):
    print(f"[Info] {msg}", file=sys.stderr, flush=True)


def logWarning(msg: str):
    print(f"[Warning] {msg}", file=sys.stderr, flush=True)


def prepareOutputFile(path: str):
    with hdf5.File(path, "a") as f:
        for g in (basisPath, hamiltonianPath, observablesPath, workspacePath):
            f.create_group(g)


def writeDataset(f: hdf5.File, group: str, name: str, x):
    if f.exists(f"{group}/{name}"):
        logWarning(f"Overwriting {group}/{name}...")
        f.delete(f"{group}/{name}")
    f.write_dataset(f"{group}/{name}", x)


def buildRepresentatives(uc: C.UserConfig):
    representatives = None
    with hdf5.File(uc.cOutput, "a") as f:
        if f.exists(f"{basisPath}/representatives"):
            logInfo(f"Loading representatives from {basisPath}/representatives...")
            representatives = f.read_dataset(f"{basisPath}/representatives")
    if representatives is None:
        logInfo("Building a list of representatives...")
    ffi.buildBasis(uc.cBasis, representatives)
    logInfo(f"Hilbert space dimension is {ffi.getNumberStates(uc.cBasis)}")
    if representatives is None:
        with hdf5.File(uc.cOutput, "a") as f:
            f.write_dataset(f"{basisPath}/representatives", ffi.basisGetStates(uc.cBasis))


def withDatatype(isReal: bool, datatype: str) -> np.dtype:
    return np.dtype({(True, "float32"): np.float32, (True, "float64"): np.float64,
                     (False, "float32"): np.complex64, (False, "float64"): np.complex128}[(isReal, datatype)])


def diagonalize(uc: C.UserConfig, dtype: np.dtype):
    op = uc.cHamiltonian.operatorObject
    logInfo(f"Diagonalizing {uc.cHamiltonian.operatorName}...")

    def monitor(info):
        logInfo(f"iteration {info['iteration']}: basis {info['basis_size']}, converged {info['number_converged']}, "
                f"matvecs {info['number_matvecs']}, evals {info['evals']}, rnorms {info['rnorms']}")
        return False

    evals, evecs, rnorms = ffi.eigh(op, dtype, uc.cNumEvals, uc.cEps, uc.cMaxBasisSize, uc.cMaxBlockSize,
                                    uc.cMinRestartSize, monitor=monitor if ffi.lib() and _debug else None)
    logInfo(f"Obtained eigenvalues {list(evals)}")
    real = np.float32 if dtype.itemsize // (2 if dtype.kind == 'c' else 1) == 4 else np.float64
    with hdf5.File(uc.cOutput, "a") as f:
        writeDataset(f, hamiltonianPath, "eigenvalues", evals.astype(real))
        # Block (N, k) column-major is written as a row-major (k, N) matrix (SpinED.hs:362-368)
        writeDataset(f, hamiltonianPath, "eigenvectors", np.ascontiguousarray(evecs.T))
        writeDataset(f, hamiltonianPath, "residuals", rnorms.astype(real))
    return evals, evecs, rnorms


def computeExpectations(uc: C.UserConfig, evecs):
    for o in uc.cObservables:
        logInfo(f"Computing expectation values of {o.operatorName}...")
        m = ffi.expectation(o.operatorObject, evecs)
        logInfo(f"Obtained expectation values {list(m)}")
        with hdf5.File(uc.cOutput, "a") as f:
            writeDataset(f, observablesPath, o.operatorName, m)


_debug = False


def run(spec: C.ConfigSpec, debug: bool = False):
    global _debug
    _debug = debug
    if debug:
        ffi.lib().ls_enable_logging()
    t0 = time.time()
    uc = C.toConfig(spec, log=logInfo)
    prepareOutputFile(uc.cOutput)
    buildRepresentatives(uc)
    isReal = ffi.isOperatorReal(uc.cHamiltonian.operatorObject)
    dtype = withDatatype(isReal, uc.cDatatype)
    evals, evecs, rnorms = diagonalize(uc, dtype)
    computeExpectations(uc, evecs)
    logInfo(f"done in {time.time() - t0:.2f} s")
    return evals, evecs, rnorms


def main(argv=None):
    ap = argparse.ArgumentParser(prog="spin-ed", description="exact diagonalization of spin systems (B200 back end)")
    ap.add_argument("--debug", action="store_true", help="Enable debug output from the library")
    ap.add_argument("input_file", help="Input yaml file")
    a = ap.parse_args(argv)
    run(C.readConfig(a.input_file), a.debug)


if __name__ == "__main__":
    main()
