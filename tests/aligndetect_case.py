"""Seeded scenario behind tests/golden/align_detect.json, shared by the fixture generator
(tests/golden/make_aligndetect_golden.py, which runs the reference on it) and by tests/test_align_detect.py."""
import base64
import json
import zlib
from pathlib import Path

import numpy as np

from ampliconreconstructorom_b200 import synth


def pack(obj) -> str:
    return base64.b64encode(zlib.compress(json.dumps(obj).encode(), 9)).decode()


def unpack(blob: str):
    return json.loads(zlib.decompress(base64.b64decode(blob)))


def scenario(tmp: Path):
    rng = np.random.default_rng(303)
    chroms = {}
    for c in range(1, 25):  # 24 chromosomes like hg19: with a handful the reference's threshold window is out of range
        chroms[c] = synth.genome_positions(rng, 2600 if c <= 3 else 450, "BspQI")
    # reference genome CMAP + key file (generate_cmap.py layout: CompntId CompntName CompntLength)
    genome = {c: np.concatenate([g, [g[-1] + 5000.0]]) for c, g in chroms.items()}
    synth.write_cmap(tmp / "genome.cmap", genome)
    with open(tmp / "genome_key.txt", "w") as f:
        f.write("# CMAP = genome.cmap\n# filter: Minimum Labels = 0\nCompntId\tCompntName\tCompntLength\n")
        for c in chroms:
            f.write(f"{c}\tchr{c}\t{genome[c][-1]:.1f}\n")
    # graph segments from chr1 and chr2, and for each a contig window that contains it and runs far beyond it on both
    # sides (large unaligned flanks); one contig has a flank of more than 5 Mb (cut to 500 labels by the reference).
    # >= 17 regions are needed for the detection thresholds to be defined (Lc = round(0.15 * #regions) >= 3).
    segs, contigs = {}, {}
    k = 0
    for c in (1, 2):
        for i in range(7):
            k += 1
            s0 = 150 + i * 330 + int(rng.integers(0, 30))
            n_seg = int(rng.integers(25, 60))
            segs[k] = synth.window_map(chroms[c], s0, n_seg)
            left, right = int(rng.integers(60, 130)), int(rng.integers(60, 130))
            if k == 3:
                left, right = 40, 820
            contigs[k] = synth.noisy_contig(rng, chroms[c], s0 - left, left + n_seg + right)
    for _ in range(40):  # unrelated contigs (chr3 only) so that the score thresholds are well defined
        k += 1
        n = int(rng.integers(40, 160))
        contigs[k] = synth.noisy_contig(rng, chroms[3], int(rng.integers(1, 2400)), n)
    synth.write_cmap(tmp / "segs.cmap", segs)
    synth.write_cmap(tmp / "contigs.cmap", contigs)
