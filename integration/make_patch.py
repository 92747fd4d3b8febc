#!/usr/bin/env python
"""integration/make_patch.py -- regenerates integration/spin-ed.patch.

Copies the three reference files the integration touches (configure, src/SpinED.hs,
src/SpinED/Internal.hs) from /root/reference into a scratch directory, applies the edits below and
writes the unified diff (`patch -p1` from the root of a twesterhout/spin-ed checkout).  The edits
are the whole Haskell-side integration of libsped: link line, the `sped_eigh` import, and the one
call in `diagonalize` that used to go to PRIMME.  GHC is not available in the build container, so
the patch is checked for applying cleanly (tests/test_integration_patch.py), not compiled.
"""
import os
import shutil
import subprocess
import sys
import tempfile

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = ["configure", "src/SpinED.hs", "src/SpinED/Internal.hs"]

INTERNAL_NEW = '''-- | Device-resident block Davidson solver of libsped (include/sped.h, sped_eigh).  It replaces the
-- call to PRIMME's @eigh@: the Krylov vectors never leave the GPU(s), only the converged
-- eigenvalues, residual norms and eigenvectors come back.
foreign import ccall safe "sped_eigh"
  sped_eigh ::
    Ptr () ->
    CInt ->
    Word64 ->
    CDouble ->
    CInt ->
    CInt ->
    CInt ->
    Ptr CDouble ->
    Ptr () ->
    Ptr CDouble ->
    FunPtr (Ptr () -> Ptr () -> IO CInt) ->
    Ptr () ->
    IO CInt

-- | Eigenvalues and residual norms always come back in double precision; convert them to the real
-- part type of the storage datatype so that the HDF5 datasets keep the reference's types.
fromDoubles :: forall a. BlasDatatype a => Proxy a -> Vector Double -> Vector (BlasRealPart a)
fromDoubles _ v = case blasTag (Proxy @a) of
  FloatTag -> V.map realToFrac v
  DoubleTag -> v
  ComplexFloatTag -> V.map realToFrac v
  ComplexDoubleTag -> v

-- | @eighDevice op dim k eps maxBasis maxBlock minRestart@: the @k@ smallest eigenpairs.
-- A size of 0 lets the solver choose (the meaning of PRIMME's defaults).
eighDevice ::
  forall a.
  BlasDatatype a =>
  Operator' ->
  Int ->
  Int ->
  Double ->
  Int ->
  Int ->
  Int ->
  IO (Vector (BlasRealPart a), Block a, Vector (BlasRealPart a))
eighDevice (Operator' op) dim k eps maxBasis maxBlock minRestart = do
  (evals :: MVector RealWorld Double) <- MV.new k
  (rnorms :: MVector RealWorld Double) <- MV.new k
  (evecs :: MVector RealWorld a) <- MV.new (dim * k)
  withForeignPtr op $ \\opPtr ->
    MV.unsafeWith evals $ \\evalsPtr ->
      MV.unsafeWith rnorms $ \\rnormsPtr ->
        MV.unsafeWith evecs $ \\evecsPtr ->
          checkStatus
            =<< sped_eigh
              opPtr
              (toCdatatype . blasTag $ Proxy @a)
              (fromIntegral k)
              (coerce eps)
              (fromIntegral maxBasis)
              (fromIntegral maxBlock)
              (fromIntegral minRestart)
              (castPtr evalsPtr)
              (castPtr evecsPtr)
              (castPtr rnormsPtr)
              nullFunPtr
              nullPtr
  evals' <- fromDoubles (Proxy @a) <$> V.unsafeFreeze evals
  rnorms' <- fromDoubles (Proxy @a) <$> V.unsafeFreeze rnorms
  evecs' <- Block (dim, k) dim <$> V.unsafeFreeze evecs
  return (evals', evecs', rnorms')

'''

DIAGONALIZE_NEW = '''  -- Sizes the input leaves unset are passed as 0: libsped then picks basis, block and restart
  -- sizes itself (what PRIMME's finalizeOptions did for its own defaults).
  let orDefault n = if n > 0 then n else 0
      maxBasis = orDefault (cMaxBasisSize config)
      maxBlock = orDefault (cMaxBlockSize config)
      minRestart = orDefault (cMinRestartSize config)
  logDebug $
    "Using max_primme_basis_size=" <> show maxBasis
      <> ", max_primme_block_size="
      <> show maxBlock
      <> ", min_primme_restart_size="
      <> show minRestart
  let hamiltonian = cHamiltonian config
  logInfo $ "Diagonalizing " <> operatorName hamiltonian <> " on the GPU..."
  result@(evals, evecs, rnorms) <-
    liftIO $
      eighDevice
        (operatorObject hamiltonian)
        dim
        (cNumEvals config)
        (cEps config)
        maxBasis
        maxBlock
        minRestart
'''

BUILDINFO_NEW = '''build_libsped() {
  # nvcc -gencode arch=compute_100a,code=sm_100a for every kernel; see csrc/Makefile
  make -C "${SPED_ROOT}/spin-ed_b200/csrc" -j"$(get_num_procs)"
}

generate_buildinfo() {
  cat <<-EOF
include-dirs:
    ${SPED_ROOT}/include
extra-lib-dirs:
    ${SPED_ROOT}/spin-ed_b200/lib
extra-libraries:
    sped
ld-options:
    -Wl,-rpath,${SPED_ROOT}/spin-ed_b200/lib
EOF
}'''


def between(text, start, stop):
    a = text.index(start)
    return a, text.index(stop, a)


def edit(path, rel):
    s = open(path).read()
    if rel == "src/SpinED/Internal.hs":
        s = s.replace("-- #include <lattice_symmetries/lattice_symmetries.h>",
                      "-- #include <sped.h>   (libsped exports the ls_* symbols of lattice_symmetries.h unchanged, plus sped_*)")
        s = s.replace("isOperatorReal' :: Operator' -> IO Bool", INTERNAL_NEW + "isOperatorReal' :: Operator' -> IO Bool")
    elif rel == "src/SpinED.hs":
        a, b = between(s, "  runLoggerInIO <- askRunInIO\n  let primmeMonitor", '  logInfo $ "Obtained eigenvalues " <> show evals')
        s = s[:a] + DIAGONALIZE_NEW + s[b:]
    else:
        s = s.replace('PREFIX="${SCRIPT_DIR}/third_party/lattice-symmetries"\n',
                      'PREFIX="${SCRIPT_DIR}/third_party/lattice-symmetries"\n'
                      "# libsped: B200 back end exporting the ls_* C API (SPED_ROOT = checkout of the sped-b200 repository)\n"
                      'SPED_ROOT="${SPED_ROOT:-${SCRIPT_DIR}/third_party/sped-b200}"\n')
        a, b = between(s, "generate_buildinfo() {", "\nprint_help() {")
        s = s[:a] + BUILDINFO_NEW + "\n" + s[b:]
        s = s.replace('    download_lattice_symmetries\n    build_static_lib\n    generate_buildinfo >"spin-ed.buildinfo"',
                      '    build_libsped\n    generate_buildinfo >"spin-ed.buildinfo"')
    open(path, "w").write(s)


def main():
    tmp = tempfile.mkdtemp()
    try:
        for side in ("a", "b"):
            for rel in FILES:
                dst = os.path.join(tmp, side, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copy(os.path.join(REF, rel), dst)
        for rel in FILES:
            edit(os.path.join(tmp, "b", rel), rel)
        out = []
        for rel in FILES:
            r = subprocess.run(["diff", "-u", "--label", "a/" + rel, "--label", "b/" + rel, os.path.join("a", rel), os.path.join("b", rel)],
                               cwd=tmp, capture_output=True, text=True)
            assert r.returncode == 1, (rel, r.returncode, r.stderr)
            out.append(r.stdout)
        with open(os.path.join(HERE, "spin-ed.patch"), "w") as f:
            f.write("".join(out))
    finally:
        shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
