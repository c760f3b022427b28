"""Seeded synthetic optical maps shaped like the workloads of BASELINE.json.

Label intervals follow the empirical hg19 BspQI distribution (log-mean 8.279, log-sd 1.395; DLE1 = the same
log-normal rescaled to a 6 kb mean).  A contig is a window of the in-silico genome seen through an optical
mapping error model: global stretch ~N(1, 0.01), per-label jitter ~N(0, 300 bp), 10 % missed labels, one false
label per 30, positions rounded to 0.1 bp; half the contigs are reversed.  Segments are exact windows.
Everything here only PRODUCES inputs (numpy); no alignment arithmetic lives in this module.
"""
from __future__ import annotations

import numpy as np

LOG_MEAN_BSPQI, LOG_SD = 8.279, 1.395


def genome_positions(rng: np.random.Generator, n_labels: int, enzyme: str = "BspQI") -> np.ndarray:
    iv = rng.lognormal(LOG_MEAN_BSPQI, LOG_SD, size=n_labels)
    if enzyme == "DLE1":
        iv *= 6000.0 / 8058.0
    iv = np.maximum(np.round(iv), 20.0)
    return np.cumsum(iv)


def window_map(genome: np.ndarray, start: int, n: int) -> np.ndarray:
    """Exact in-silico map of n labels: [pos_0..pos_{n-1}, length] with a flank either side."""
    w = genome[start:start + n]
    left = w[0] - (genome[start - 1] if start > 0 else 0.0)
    p = w - w[0] + np.floor(left / 2) + 1
    right = (genome[start + n] - w[-1]) if start + n < genome.size else 2000.0
    length = p[-1] + np.floor(right / 2) + 1
    return np.concatenate([p, [length]]).astype(np.float64)


def noisy_contig(rng: np.random.Generator, genome: np.ndarray, start: int, n: int) -> np.ndarray:
    m = window_map(genome, start, n)
    p, length = m[:-1], m[-1]
    keep = rng.random(p.size) > 0.10
    if keep.sum() < 2:
        keep[:2] = True
    p = p[keep]
    n_extra = rng.binomial(p.size, 1.0 / 30.0)
    if n_extra:
        p = np.concatenate([p, rng.uniform(1.0, length - 1.0, size=n_extra)])
    scale = rng.normal(1.0, 0.01)
    p = np.sort(p * scale + rng.normal(0.0, 300.0, size=p.size))
    p = p - min(p[0], 0.0) + (20.0 if p[0] <= 0 else 0.0)
    length = max(length * scale, p[-1] + 20.0)
    if rng.random() < 0.5:
        p = (length - p)[::-1]
    out = np.round(np.concatenate([p, [length]]), 1)
    # strictly increasing after rounding
    for k in range(1, out.size):
        if out[k] <= out[k - 1]:
            out[k] = out[k - 1] + 0.1
    return np.round(out, 1)


def lognormal_sizes(rng, n, median, sigma, lo, hi):
    return np.clip(np.round(rng.lognormal(np.log(median), sigma, size=n)), lo, hi).astype(np.int64)


def make_workload(seed: int, n_segs: int, n_contigs: int, seg_median=40, seg_sigma=0.8, seg_clip=(11, 500),
                  contig_median=2500, contig_sigma=0.6, contig_clip=(500, 8000), enzyme="DLE1",
                  genome_labels: int | None = None):
    """C2-shaped workload (defaults = SURVEY 8(d) C2): returns (segs: {id: map}, contigs: {id: map})."""
    rng = np.random.default_rng(seed)
    nx = lognormal_sizes(rng, n_segs, seg_median, seg_sigma, *seg_clip)
    nb = lognormal_sizes(rng, n_contigs, contig_median, contig_sigma, *contig_clip)
    if genome_labels is None:
        genome_labels = int(max(4 * nb.max(), 100_000))
    g = genome_positions(rng, genome_labels, enzyme)
    segs, contigs = {}, {}
    for k in range(n_segs):
        s = int(rng.integers(1, genome_labels - nx[k] - 1))
        segs[k + 1] = window_map(g, s, int(nx[k]))
    for k in range(n_contigs):
        s = int(rng.integers(1, genome_labels - nb[k] - 1))
        contigs[k + 1] = noisy_contig(rng, g, s, int(nb[k]))
    return segs, contigs


def write_cmap(path, maps: dict) -> None:
    """CMAP text in the column layout parse_cmap reads (col 0 = CMapId, col 5 = Position; SAHelper.cpp:30-40)."""
    with open(path, "w") as f:
        f.write("# CMAP File Version:\t0.1\n# Label Channels:\t1\n")
        f.write(f"# Number of Consensus Nanomaps:\t{len(maps)}\n")
        f.write("#h CMapId\tContigLength\tNumSites\tSiteID\tLabelChannel\tPosition\tStdDev\tCoverage\tOccurrence\n")
        f.write("#f int\tfloat\tint\tint\tint\tfloat\tfloat\tint\tint\n")
        for mid in sorted(maps):
            m = maps[mid]
            n = len(m) - 1
            length = f"{m[-1]:.1f}"
            for k, p in enumerate(m):
                ch = "1" if k < n else "0"
                f.write(f"{mid}\t{length}\t{n}\t{k + 1}\t{ch}\t{p:.1f}\t1.0\t1\t1\n")


def flatten(maps: dict):
    """{id: array} -> (ids, off, pos float32) in ascending-id order (std::map order of the reference)."""
    ids = np.array(sorted(maps), dtype=np.int32)
    off = np.zeros(len(ids) + 1, dtype=np.int64)
    for k, i in enumerate(ids):
        off[k + 1] = off[k] + len(maps[i])
    pos = np.concatenate([np.asarray(maps[i], dtype=np.float64) for i in ids]).astype(np.float32)
    return ids, off, pos
