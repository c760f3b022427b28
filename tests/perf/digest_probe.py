"""Developer probe: one digestion of 1 GiB of random sequence through the C ABI (for ncu captures of digest_kernel)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from ampliconreconstructorom_b200 import capi
ctx = capi.Context(0)
rng = np.random.default_rng(1)
seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=1 << 30, dtype=np.uint8)]
for rep in range(3):
    pos = ctx.digest(seq, "GCTCTTC", 9)
    ms = ctx.timing().fill_ms
    print(f"rep{rep}: {seq.size/1e6:.0f} MB in {ms:.3f} ms = {seq.size/ms/1e6:.0f} GB/s; {pos.size} labels")
