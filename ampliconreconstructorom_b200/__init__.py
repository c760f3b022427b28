"""B200-native SegAligner (the DP hot path of AmpliconSuite/AmpliconReconstructorOM).

Layout: csrc/ = sm_100a kernels + C ABI + C++ host (the drop-in `SegAligner` executable); capi.py = ctypes
binding used by tests and bench.py; synth.py = seeded synthetic CMAP workloads; segaligner.py = thin wrapper
that runs the executable the way ARAlignDetect.py / OMPathFinder.py do.
"""
from . import capi, synth  # noqa: F401

__all__ = ["capi", "synth"]
