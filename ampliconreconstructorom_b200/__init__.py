"""B200-native SegAligner (the DP hot path of AmpliconSuite/AmpliconReconstructorOM).

Layout: csrc/ = sm_100a kernels + C ABI + C++ host (the drop-in `SegAligner` executable); capi.py = ctypes
binding used by tests and bench.py; synth.py = seeded synthetic CMAP workloads; generate_cmap.py = drop-in in-silico digestion
(GPU motif scan); align_detect.py = drop-in region extraction / detection post-processing of ARAlignDetect.py; dist_util.py =
sharding helpers of bench.py.
"""
from . import capi, synth  # noqa: F401

__all__ = ["capi", "synth"]
