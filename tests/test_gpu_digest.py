"""GPU: in-silico digestion.  The drop-in generate_cmap against digests of the UNMODIFIED
reference generate_cmap.py recorded in the build container (tests/golden/digest_digests.json, inputs regenerated
from seeds), and the sa_digest C-ABI entry against the reference's own formulation (two look-ahead regex scans)."""
import hashlib
import json
import re
import subprocess
import sys

import numpy as np
import pytest

import digest_cases as dc
import helpers
from ampliconreconstructorom_b200 import capi

pytestmark = pytest.mark.gpu
GOLD = json.loads((helpers.GOLDEN / "digest_digests.json").read_text())
COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def ref_positions(seq: str, motif: str, table_offset: int):
    """generate_cmap.py:102-123 verbatim (the regex formulation), as the oracle for the kernel."""
    offset = table_offset + 1
    seq = seq.upper()
    rev = "".join(COMP[a] for a in motif[::-1])
    rev_offset = len(motif) - offset
    pos = [float(m.start() + offset) for m in re.finditer("(?=" + motif + ")", seq)]
    neg = [float(m.start()) + rev_offset for m in re.finditer("(?=" + rev + ")", seq)]
    return sorted(list(set(neg + pos)))


@pytest.mark.parametrize("enzyme,motif,off", [("BSPQI", "GCTCTTC", 8), ("BBVCI", "CCTCAGC", 2), ("BSMI", "GAATGC", 4),
                                               ("BSRDI", "GCAATG", 5), ("DLE1", "CTTAAG", 2)])
def test_sa_digest_matches_regex_formulation(ctx, enzyme, motif, off):
    rng = np.random.default_rng(len(motif) * 100 + off)
    for n in (0, 1, 5, 6, 7, 15, 16, 17, 31, 33, 1000, 65_537, 300_001):
        seq = "".join(rng.choice(list("ACGTacgtN"), size=n, p=[.22, .22, .22, .22, .02, .02, .02, .02, .04]).tolist())
        if n >= 40:  # overlapping, adjacent and boundary placements on both strands
            rev = "".join(COMP[a] for a in motif[::-1])
            seq = motif + rev + seq[2 * len(motif):n - 2 * len(motif)] + rev + motif
            k = n // 2
            seq = seq[:k] + motif + motif[1:] + seq[k + 2 * len(motif) - 1:]
            seq = seq[:n]
        got = ctx.digest(seq.encode(), motif, off + 1)
        exp = ref_positions(seq, motif, off)
        assert got.tolist() == [int(v) for v in exp], f"{enzyme} n={n}"


@pytest.mark.parametrize("name", sorted(dc.CASES))
def test_generate_cmap_matches_reference_digests(tmp_path, name):
    dc.write_inputs(tmp_path)
    script = helpers.ROOT / "ampliconreconstructorom_b200" / "generate_cmap.py"
    p = subprocess.run([sys.executable, str(script), "-r", "ref.fa", "-o", "out"] + dc.CASES[name], cwd=tmp_path,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    got = {f.name: hashlib.sha256(f.read_bytes()).hexdigest() for f in sorted(tmp_path.iterdir()) if f.name.startswith("out")}
    got["stdout"] = hashlib.sha256(p.stdout.encode()).hexdigest()
    assert got == GOLD[name]


def test_digest_throughput_report(ctx):
    """Not a pass/fail performance gate: records the kernel's scan rate on 1 GiB of sequence (8x the L2)."""
    rng = np.random.default_rng(1)
    seq = rng.integers(0, 4, size=1 << 30, dtype=np.uint8)
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[seq]
    pos = ctx.digest(seq, "GCTCTTC", 9)
    ms = ctx.timing().fill_ms
    print(f"digest: {seq.size / 1e6:.0f} MB scanned in {ms:.2f} ms kernel time = {seq.size / ms / 1e6:.0f} GB/s, {pos.size} labels")
    assert pos.size > 10_000 and np.all(np.diff(pos) > 0)
