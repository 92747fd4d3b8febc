"""sped-b200: B200-native back end for SpinED's hot path (representative basis + symmetry-adapted
matrix-free matvec + ground-state eigensolve).  The product is ``lib/libsped.so`` (C ABI in
``include/sped.h``); this package is the Python stand-in for the reference's Haskell host layers
(``src/SpinED/Internal.hs`` -> ``ffi``, ``src/SpinED.hs`` -> ``config`` / ``driver``)."""
__version__ = "0.1.0"
