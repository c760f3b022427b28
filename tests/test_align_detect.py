"""CPU: SURVEY.md section 8 row f-3 -- ampliconreconstructorom_b200/align_detect.py against the reference's own
ARAlignDetect.py functions (:31-117, :137-155, :198-280).  The fixture tests/golden/align_detect.json was produced by
tests/golden/make_aligndetect_golden.py, which IMPORTS the unmodified reference module and runs the unmodified reference
SegAligner on the seeded scenario of tests/aligndetect_case.py; every returned structure and every written byte must match."""
import hashlib
import io
import json
import os
from contextlib import redirect_stdout
from pathlib import Path

import pytest

import aligndetect_case as case
from ampliconreconstructorom_b200 import align_detect as ad

GOLD = Path(__file__).resolve().parent / "golden" / "align_detect.json"


@pytest.fixture(scope="module")
def gold():
    return case.unpack(json.loads(GOLD.read_text())["blob"])


@pytest.fixture()
def work(tmp_path, gold):
    case.scenario(tmp_path)
    for n, h in gold["input_sha256"].items():
        assert hashlib.sha256((tmp_path / n).read_bytes()).hexdigest() == h, f"regenerated {n} differs from the fixture's input"
    a_dir = tmp_path / "alignments"
    a_dir.mkdir()
    for n, txt in list(gold["aln_files"].items()) + list(gold["det_files"].items()):
        (a_dir / n).write_text(txt)
    return tmp_path


def test_unaligned_regions_and_region_cmap(work, gold):
    cm = ad.parse_cmap(str(work / "contigs.cmap"))
    cl = ad.get_cmap_lens(str(work / "contigs.cmap"))
    regions = ad.get_unaligned_segs(str(work / "alignments") + "/", gold["aln_flist"], contig_cmaps=cm)
    assert list(regions.keys()) == gold["regions_order"]
    assert {c: v for c, v in regions.items()} == gold["regions"]
    assert any(len(l) == ad.search_upper_cutoff_labels or len(l) > 200 for v in regions.values() for l in v)
    trans, fname, cid_d = ad.write_unaligned_cmaps(regions, str(work) + "/", "BspQI", contig_cmaps=cm, contig_lens=cl)
    assert fname == str(work) + "/contig_unaligned_regions.cmap"
    assert Path(fname).read_text() == gold["regions_cmap"]
    assert trans == gold["label_trans"] and cid_d == gold["cid_d"]
    # the module-global form the reference uses (ARAlignDetect.py:359-361)
    ad.contig_cmaps, ad.contig_lens = cm, cl
    try:
        r2 = ad.get_unaligned_segs(str(work / "alignments") + "/", gold["aln_flist"])
        os.remove(fname)
        t2, f2, c2 = ad.write_unaligned_cmaps(r2, str(work) + "/", "BspQI")
        assert Path(f2).read_text() == gold["regions_cmap"] and t2 == trans and c2 == cid_d
    finally:
        ad.contig_cmaps, ad.contig_lens = {}, {}


def test_aligned_labels_file(work, gold):
    name = ad.write_aligned_labels(str(work / "alignments") + "/", gold["aln_flist"], str(work) + "/")
    assert Path(name).read_text() == gold["aligned_labels"]


def test_detections_to_graph_segments(work, gold):
    a_dir = work / "alignments"
    before = set(os.listdir(a_dir))
    buf = io.StringIO()
    with redirect_stdout(buf):
        bpg = ad.detections_to_seg_alignments(str(a_dir) + "/", gold["det_flist"], str(work / "genome.cmap"), gold["cid_d"],
                                              gold["label_trans"], 14, "SA_segs")
    assert [list(x) for x in bpg] == gold["bpg_list"]
    assert buf.getvalue() == gold["det_stdout"]
    new = sorted(set(os.listdir(a_dir)) - before)
    assert new == sorted(gold["rg_files"])
    for n in new:
        assert (a_dir / n).read_text() == gold["rg_files"][n], n


def test_region_extraction_edge_cases(tmp_path):
    """No free labels, > 90 % unaligned, a run that ends at the last label (map end falls back to contig_lens), runs at
    the size / label-count cut-offs.  The expectations below were checked against the reference's own functions on the
    same inputs when the test was written (the reference module needs /root/reference, absent at test time)."""
    cmap = tmp_path / "c.cmap"
    n = 60
    with open(cmap, "w") as f:
        f.write("#h CMapId\tContigLength\tNumSites\tSiteID\tLabelChannel\tPosition\tStdDev\tCoverage\tOccurrence\n")
        for cid in ("1", "2", "3"):
            for k in range(n):
                f.write(f"{cid}\t{n * 10000.0 + 500.0}\t{n}\t{k + 1}\t1\t{(k + 1) * 10000.0}\t1.0\t1\t1\n")
            f.write(f"{cid}\t{n * 10000.0 + 500.0}\t{n}\t{n + 1}\t0\t{n * 10000.0 + 500.0}\t1.0\t1\t1\n")
    cm, cl = ad.parse_cmap(str(cmap)), ad.get_cmap_lens(str(cmap))

    def aln(name, contig, a, b):
        (tmp_path / name).write_text("#seg_seq\ttotal_score\tcircular\n#1+\t1000\tFalse\n"
                                     "#contig_id\tseg_id\tcontig_label\tseg_label\tcontig_dir\tseg_dir\tseg_aln_number\tscore\tscore_delta\n"
                                     f"{contig}\t1\t{a}\t1\t+\t+\t0\t0\t0\n{contig}\t1\t{b}\t9\t+\t+\t0\t1000\t1000\n")
    aln("a_aln.txt", "1", 1, 30)     # contig 1: labels 30.. free would be index 31..59 (29 labels > 20, 280 kb > 200 kb)
    aln("b_aln.txt", "2", 1, 60)     # contig 2: only index 0 stays free -> too short
    aln("c_aln.txt", "3", 58, 60)    # contig 3: 95 % unaligned -> skipped
    reg = ad.get_unaligned_segs(str(tmp_path) + "/", ["a_aln.txt", "b_aln.txt", "c_aln.txt"], contig_cmaps=cm)
    assert dict(reg) == {"1": [list(range(31, 60))]}
    trans, fname, cid = ad.write_unaligned_cmaps(reg, str(tmp_path) + "/", "DLE1", contig_cmaps=cm, contig_lens=cl)
    rows = [l.split("\t") for l in Path(fname).read_text().splitlines() if not l.startswith("#")]
    assert rows[0][:6] == ["4", str(cl["1"] - 320000.0), "29", "1", "1", "0.0"]
    assert rows[-1][3:6] == ["29", "0", str(cl["1"] - 320000.0)]   # end row: contig length (no label l[-1]+2), SiteID repeats
    assert trans["4"]["1"] == "32" and cid == {"4": "1"}
