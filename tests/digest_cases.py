"""Seeded FASTA / graph / BED inputs for the in-silico digestion drop-in (generate_cmap), shared by the golden
generator (reference run in the build container) and the GPU test."""
from pathlib import Path

import numpy as np

MOTIFS = ["GCTCTTC", "GAAGAGC", "CTTAAG", "CCTCAGC", "GCTGAGG", "GAATGC", "GCATTC", "GCAATG", "CATTGC"]


def write_inputs(d: Path, seed: int = 7):
    rng = np.random.default_rng(seed)
    d.mkdir(parents=True, exist_ok=True)
    names = ["chr1", "chr2", "chr10", "chrX", "scaffold_7 extra words"]
    sizes = [180_000, 95_000, 61_000, 40_003, 1_500]
    with open(d / "ref.fa", "w") as f:
        for name, n in zip(names, sizes):
            seq = rng.choice(list("ACGT"), size=n, p=[0.3, 0.2, 0.2, 0.3])
            seq = "".join(seq.tolist())
            seq = list(seq)
            for k in range(n // 4000):  # plant motifs incl. overlapping / back-to-back / edge placements
                p = int(rng.integers(0, n - 16))
                m = MOTIFS[int(rng.integers(0, len(MOTIFS)))]
                seq[p:p + len(m)] = list(m)
            seq[0:7] = list("GAAGAGC")      # reverse-strand BspQI site at the very start -> negative label position
            seq[n - 7:n] = list("GCTCTTC")  # forward site flush with the end -> position beyond the map
            seq = "".join(seq)
            if n > 10_000:
                lo = int(rng.integers(100, n - 5000))
                seq = seq[:lo] + seq[lo:lo + 3000].lower() + "N" * 200 + seq[lo + 3200:]  # soft-masked + N run
            f.write(">" + name + "\n")
            w = 60 if name != "chr2" else 71
            for k in range(0, len(seq), w):
                f.write(seq[k:k + w] + "\n")
    with open(d / "graph.txt", "w") as f:
        f.write("SequenceEdge: StartPosition, EndPosition, PredictedCopyCount, AverageCoverage, Size, NumberReadsMapped\n")
        f.write("sequence\tchr1:1000-\tchr1:61000+\t4.0\t30.0\t60001\t1000\n")
        f.write("sequence\tchr2:5000-\tchr2:25000+\t4.0\t30.0\t20001\t1000\n")
        f.write("sequence\tchrX:39000-\tchrX:40003+\t2.0\t30.0\t1004\t100\n")
        f.write("concordant\tchr1:61000+->chr2:5000-\t3\n")
    with open(d / "regions.bed", "w") as f:
        f.write("#chrom\tstart\tend\n")
        f.write("chr10\t100\t50000\nchr10\t50001\t61000\nchr1\t150000\t179999\n")


CASES = {
    "ref_bspqi": ["--makeRef", "-e", "BspQI"],
    "ref_dle1_filtered": ["--makeRef", "-e", "dle1", "-l", "3", "-s", "2000"],
    "graph_bspqi": ["-g", "graph.txt", "-e", "BspQI", "-l", "1"],
    "bed_bbvci": ["-b", "regions.bed", "-e", "BbvCI"],
    "ref_bsmi": ["--makeRef", "-e", "BsmI", "-s", "50000"],
    "ref_bsrdi": ["--makeRef", "-e", "BsrDI"],
}
