"""CPU: host-side product code (csrc/host/cmap.cpp through the C ABI) -- CMAP parsing quirks, thresholds,
number formatting -- against the reference (when oracle/_ref is built) and against fixed expectations."""
import numpy as np
import pytest

import helpers
from ampliconreconstructorom_b200 import capi, synth


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


CMAP_QUIRKS = (
    "# header\n"
    "#h CMapId\tContigLength\tNumSites\tSiteID\tLabelChannel\tPosition\tStdDev\tCoverage\tOccurrence\n"
    "7\t1000.5\t2\t1\t1\t100.25\t1.0\t1\t1\n"
    "7\t1000.5\t2\t2\t1\t500.125\t1.0\t1\t1\n"
    "7\t1000.5\t2\t3\t0\t1000.5\t1.0\t1\t1\n"
    "3\t2000.0\t1\t1\t1\t1234567.89\t1.0\t1\t1\n"
    "\n"                                           # blank line: pushes the stale position again (SAHelper.cpp:44)
    "3\t2000.0\t1\t2\t0\t2000.0\n"                 # position not followed by a tab: not parsed, stale value pushed
    "9\tshort\n"                                   # too few columns: new map 9 gets the stale position
    "7\t1000.5\t2\t1\t1\t42.0\t1.0\t1\t1\n"        # id 7 re-appears: the map is re-created empty (:35)
)


def test_parse_cmap_quirks(tmp_path):
    f = tmp_path / "q.cmap"
    f.write_text(CMAP_QUIRKS)
    ids, off, pos = capi.host_parse_cmap(f)
    maps = capi.unflatten_maps(ids, off, pos)
    assert sorted(maps) == [3, 7, 9]
    assert maps[7].tolist() == [42.0]
    assert np.array_equal(bits(maps[3]), bits(np.array([1234567.89, 1234567.89, 1234567.89], dtype=np.float32)))
    assert np.array_equal(bits(maps[9]), bits(np.array([1234567.89], dtype=np.float32)))


def test_parse_cmap_matches_reference(tmp_path, ref):
    f = tmp_path / "q.cmap"
    f.write_text(CMAP_QUIRKS)
    ours = capi.unflatten_maps(*capi.host_parse_cmap(f))
    theirs = ref.parse_cmap(f)
    assert sorted(ours) == sorted(theirs) and all(np.array_equal(bits(ours[k]), bits(theirs[k])) for k in ours)
    segs, contigs = synth.make_workload(9, 6, 9, contig_median=120, contig_clip=(20, 300))
    synth.write_cmap(tmp_path / "c.cmap", contigs)
    ours = capi.unflatten_maps(*capi.host_parse_cmap(tmp_path / "c.cmap"))
    theirs = ref.parse_cmap(tmp_path / "c.cmap")
    assert sorted(ours) == sorted(theirs) and all(np.array_equal(bits(ours[k]), bits(theirs[k])) for k in ours)


def test_shipped_gbm39_cmap_parses_like_the_survey_says():
    p = helpers.ORACLE_DIR / "_ref" / "data" / "GBM39_amplicon1_graph_BSPQI.cmap"
    if not p.exists():
        pytest.skip("oracle/_ref/data not present")
    ids, off, pos = capi.host_parse_cmap(p)
    assert (np.diff(off) - 1).tolist() == [13, 54, 3, 2, 66, 0, 4, 45, 9]
    ri, _, _ = capi.host_make_reverse(ids, off, pos, 10)
    assert ri.tolist() == [-8, -5, -2, -1, 1, 2, 5, 8]  # KA-5


def _score_rows(rng, segs, contigs, per_pair=1):
    rs, rc, sc = [], [], []
    for s in sorted(segs):
        for c in sorted(contigs):
            for _ in range(per_pair):
                rs.append(s)
                rc.append(c)
                sc.append(float(np.float32(rng.normal(20000, 15000))))
    return np.array(rs), np.array(rc), np.array(sc, dtype=np.float32)


@pytest.mark.parametrize("n_contigs,n_detect,per_pair", [(60, 500, 1), (200, 500, 1), (45, 30, 3), (10, 500, 1)])
def test_score_thresholds(oracle, n_contigs, n_detect, per_pair):
    rng = np.random.default_rng(n_contigs)
    segs = {k: helpers.random_map(rng, int(rng.integers(6, 40))) for k in (-2, -1, 1, 2)}
    contigs = {k + 1: helpers.random_map(rng, int(rng.integers(20, 200))) for k in range(n_contigs)}
    rs, rc, sc = _score_rows(rng, segs, contigs, per_pair)
    for pv in (float(np.float32(1e-4)), float(np.float32(1e-6)), float(np.float32(1e-9))):
        ours = capi.host_score_thresholds(rs, rc, sc, segs, contigs, pv, n_detect)
        m = sum(len(v) - 1 for v in contigs.values())
        for s in segs:
            exp = oracle.threshold(sc[rs == s], n_contigs, n_detect, m, len(segs[s]) - 1, pv)
            assert (np.isnan(exp) and np.isnan(ours[s])) or bits(exp) == bits(ours[s])
        if n_contigs == 10:  # KA-3: too few contigs -> "-nan"
            assert all(np.isnan(v) for v in ours.values())
            assert capi.host_format_float(ours[1]) == "-nan"


def test_score_thresholds_threaded_partial_sort(oracle):
    """Cohort-sized score vectors (> 65 536 rows) take the multi-threaded path of compute_score_thresholds, which orders
    only the tail of each segment's scores that the regression reads; it must still equal the oracle, which sorts
    everything like the reference (main.cpp:59)."""
    rng = np.random.default_rng(99)
    segs = {k: helpers.random_map(rng, int(rng.integers(8, 40))) for k in list(range(-30, 0)) + list(range(1, 31))}
    n_contigs = 1500
    contigs = {k + 1: helpers.random_map(rng, int(rng.integers(20, 60))) for k in range(n_contigs)}
    s_ids, c_ids = np.array(sorted(segs)), np.array(sorted(contigs))
    rs = np.repeat(s_ids, n_contigs)
    rc = np.tile(c_ids, s_ids.size)
    sc = rng.normal(20000, 15000, size=rs.size).astype(np.float32)
    assert rs.size > 65536
    m = sum(len(v) - 1 for v in contigs.values())
    for pv in (float(np.float32(1e-4)), float(np.float32(1e-6))):
        ours = capi.host_score_thresholds(rs, rc, sc, segs, contigs, pv, 500)
        for s in segs:
            exp = oracle.threshold(sc[rs == s], n_contigs, 500, m, len(segs[s]) - 1, pv)
            assert (np.isnan(exp) and np.isnan(ours[s])) or bits(exp) == bits(ours[s])


@pytest.mark.parametrize("n_contigs,seg_lo", [(80, 12), (200, 6)])
def test_score_thresholds_match_reference(ref, n_contigs, seg_lo):
    """Segments with < 12 labels shift the regression window down by 25 (main.cpp:80-82); with fewer than 170
    contigs that makes the reference's window negative and it dies with std::length_error, so the small-segment
    case is only comparable with >= 170 contigs (ours returns -nan there instead of aborting)."""
    rng = np.random.default_rng(3)
    segs = {k: helpers.random_map(rng, int(rng.integers(seg_lo, 40))) for k in (-2, -1, 1, 2)}
    contigs = {k + 1: helpers.random_map(rng, int(rng.integers(20, 200))) for k in range(n_contigs)}
    rs, rc, sc = _score_rows(rng, segs, contigs)
    for pv in (float(np.float32(1e-4)), float(np.float32(1e-9))):
        ours = capi.host_score_thresholds(rs, rc, sc, segs, contigs, pv, 500)
        theirs = ref.thresholds(rs, rc, sc, segs, contigs, pv, 500)
        assert all(bits(ours[k]) == bits(theirs[k]) for k in ours)


def test_float_formatting_is_ostream_default():
    for v, s in [(447630.0, "447630"), (9583.512, "9583.51"), (1.33932e7, "1.33932e+07"), (0.0, "0"), (-0.5, "-0.5"),
                 (float("inf"), "inf"), (119898.0, "119898"), (1e-5, "1e-05"), (123456.7, "123457")]:
        assert capi.host_format_float(v) == s
