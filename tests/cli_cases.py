"""Shared CLI-level cases: run the reference SegAligner and the B200 SegAligner on the same CMAP files and
compare every output file byte for byte (used by the gpu-marked tests and by tests/golden/make_golden.py)."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ampliconreconstructorom_b200 import synth  # noqa: E402

REF_BIN = ROOT / "oracle" / "_ref" / "SegAligner_ref"
NEW_BIN = ROOT / "ampliconreconstructorom_b200" / "bin" / "SegAligner"


def make_case(seed: int, n_segs=6, n_contigs=60, n_related=14, seg_labels=(12, 70), contig_labels=(40, 260),
              enzyme="BspQI", integer=False):
    """Segments tile a region of an in-silico genome; `n_related` contigs are noisy windows over that region
    (true alignments, some spanning several segments, half reversed); the rest come from elsewhere (null)."""
    rng = np.random.default_rng(seed)
    g = synth.genome_positions(rng, 60_000, enzyme)
    segs, contigs = {}, {}
    start = 1000
    cur = start
    for k in range(n_segs):
        n = int(rng.integers(*seg_labels))
        segs[k + 1] = synth.window_map(g, cur, n)
        cur += n
    region_end = cur
    for k in range(n_contigs):
        n = int(rng.integers(*contig_labels))
        if k < n_related:
            s = int(rng.integers(max(start - n // 2, 1), region_end - 5))
        else:
            s = int(rng.integers(region_end + 500, g.size - n - 2))
        contigs[k + 1] = synth.noisy_contig(rng, g, s, n)
        if integer:
            contigs[k + 1] = np.round(contigs[k + 1])
    return segs, contigs


def write_case(tmp: Path, segs, contigs):
    tmp.mkdir(parents=True, exist_ok=True)
    synth.write_cmap(tmp / "segs.cmap", segs)
    synth.write_cmap(tmp / "contigs.cmap", contigs)
    return tmp / "segs.cmap", tmp / "contigs.cmap"


def run_bin(binary: Path, workdir: Path, seg_file, contig_file, flags, timeout=1800, env=None):
    workdir.mkdir(parents=True, exist_ok=True)
    cmd = [str(binary), str(seg_file), str(contig_file), f"-prefix={workdir}/SA"] + list(flags)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout, env=env)
    return p


def dir_digest(workdir: Path) -> dict:
    out = {}
    for f in sorted(workdir.iterdir()):
        if f.is_file():
            out[f.name] = hashlib.sha256(f.read_bytes()).hexdigest()
    return out


def compare_dirs(a: Path, b: Path):
    da, db = dir_digest(a), dir_digest(b)
    only_a = sorted(set(da) - set(db))
    only_b = sorted(set(db) - set(da))
    diff = sorted(k for k in set(da) & set(db) if da[k] != db[k])
    return only_a, only_b, diff, len(da)


CASES = {
    # name: (make_case kwargs, flags)
    "standard_gen2": (dict(seed=1), ["-nthreads=3", "-min_labs=10", "-gen=2"]),
    "standard_gen1": (dict(seed=2, n_contigs=70), ["-nthreads=2", "-min_labs=10", "-gen=1"]),
    "local_gen2": (dict(seed=3), ["-nthreads=4", "-min_labs=10", "-gen=2", "-local"]),
    "no_tip": (dict(seed=4), ["-nthreads=1", "-min_labs=12", "-gen=2", "-no_tip_aln"]),
    "integer_positions": (dict(seed=5, integer=True), ["-nthreads=2", "-min_labs=10", "-gen=1"]),
    "few_contigs_nan_thresholds": (dict(seed=6, n_contigs=10, n_related=4), ["-nthreads=2", "-gen=2"]),  # KA-3
    "detection": (dict(seed=7, n_segs=5, n_contigs=45, n_related=10, contig_labels=(100, 900)),
                  ["-nthreads=4", "-min_labs=10", "-gen=1", "-detection"]),
    "unbanded_nl": (dict(seed=9, n_segs=3, n_contigs=36, n_related=6, seg_labels=(12, 26), contig_labels=(30, 70)),
                    ["-nthreads=2", "-min_labs=10", "-gen=2", "-nl"]),
    "detection_small_n": (dict(seed=8, n_segs=4, n_contigs=40, n_related=8, contig_labels=(60, 400)),
                          ["-nthreads=2", "-min_labs=10", "-gen=2", "-detection", "-n_detect_scores=120"]),
}


def make_edge_case(seed: int):
    """Degenerate shapes next to normal ones: contigs with 1-3 labels, segments with 1-3 labels (kept because
    -min_labs=1), a segment that repeats a contig exactly (zero discrepancies -> the generic powf path)."""
    rng = np.random.default_rng(seed)
    # >= 170 contigs: with fewer, the reference dies with std::length_error for segments of < 12 labels (main.cpp:77-84)
    segs, contigs = make_case(seed, n_segs=3, n_contigs=176, n_related=8, seg_labels=(12, 30), contig_labels=(20, 60))
    g = synth.genome_positions(rng, 4000)
    segs[4] = synth.window_map(g, 100, 2)
    segs[5] = synth.window_map(g, 300, 3)
    segs[6] = synth.window_map(g, 500, 4)
    contigs[177] = synth.noisy_contig(rng, g, 700, 2)
    contigs[178] = synth.noisy_contig(rng, g, 900, 3)
    contigs[179] = np.array(segs[1], dtype=np.float64)  # identical to segment 1: discrepancies of exactly 0
    contigs[180] = synth.window_map(g, 100, 40)          # exact in-silico window containing segs[4]
    return segs, contigs


EDGE_FLAGS = ["-nthreads=2", "-min_labs=2", "-gen=2"]


# ---- named BASELINE configs pinned against the reference binary --------------------------------------
DATA = ROOT / "oracle" / "_ref" / "data"
C1_FLAGS = ["-min_labs=10", "-gen=1"]  # README.md:162-191 (the GBM39 test run: BspQI / Irys = -gen=1)
C5_FLAGS = ["-min_labs=12", "-gen=2"]


def make_c1(n_contigs=2500, seed=39):
    """C1: the shipped GBM39 segment CMAP (real data, argv[1]) against a seeded Irys-like BspQI assembly: `n_contigs`
    contigs with ~LogNormal(median 110, sigma 0.5) labels clipped to [15, 600] cut from an in-silico BspQI genome, plus 6
    contigs that are noisy windows of the real hg19 chr7 map over 54.5-56.5 Mb (where the GBM39 EGFR amplicon segments
    come from), so that real alignments exist.  Returns (path of the segment CMAP, {id: contig map})."""
    import gzip
    rng = np.random.default_rng(seed)
    nb = synth.lognormal_sizes(rng, n_contigs, 110, 0.5, 15, 600)
    g = synth.genome_positions(rng, 400_000, "BspQI")
    contigs = {}
    for k in range(n_contigs):
        s = int(rng.integers(1, g.size - nb[k] - 2))
        contigs[k + 1] = synth.noisy_contig(rng, g, s, int(nb[k]))
    chr7 = []
    with gzip.open(DATA / "hg19_BspQI.cmap.gz", "rt") as f:
        for line in f:
            if line.startswith("#"):
                continue
            t = line.split("\t")
            if t[0] == "7" and t[4] != "0":
                chr7.append(float(t[5]))
    chr7 = np.array(chr7)
    lo = int(np.searchsorted(chr7, 54.5e6))
    hi = int(np.searchsorted(chr7, 56.5e6))
    for k in range(6):
        n = int(rng.integers(60, 140))
        s = int(rng.integers(lo, max(hi - n, lo + 1)))
        contigs[n_contigs + k + 1] = synth.noisy_contig(rng, chr7, s, n)
    return DATA / "GBM39_amplicon1_graph_BSPQI.cmap", contigs


def make_c5_shaped(n_segs=300, n_contigs=560, seed=5):
    """A C5-shaped subsample (cohort sweep: many short graph segments x many short genome-wide DLE1 contigs) of the
    generator tests/perf/c5_cohort.py uses at full size, small enough for the reference to finish in about a minute."""
    return synth.make_workload(seed, n_segs, n_contigs, seg_median=16, seg_sigma=0.6, seg_clip=(12, 80), contig_median=80,
                               contig_sigma=0.5, contig_clip=(40, 400), genome_labels=50_000)
