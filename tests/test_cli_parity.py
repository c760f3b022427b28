"""CLI-level parity on the GPU box: the B200 `SegAligner` executable against the UNMODIFIED reference binary
(oracle/_ref/SegAligner_ref, built from /root/reference by oracle/Makefile and shipped with the snapshot).
Every output file must be byte-identical: *_all_scores.txt (same -nthreads => same row order),
*_score_thresholds.txt and every *_aln.txt, in standard, local, detection, fitting and contig-list runs."""
import numpy as np
import pytest

import cli_cases as cc
from ampliconreconstructorom_b200 import synth

pytestmark = pytest.mark.gpu


def _need_bins():
    if not cc.REF_BIN.exists():
        pytest.skip("oracle/_ref/SegAligner_ref not built")
    assert cc.NEW_BIN.exists(), "ampliconreconstructorom_b200/bin/SegAligner not built (run __graft_entry__.build())"


def _run_both(tmp_path, segs, contigs, flags):
    seg_f, con_f = cc.write_case(tmp_path / "in", segs, contigs)
    r = cc.run_bin(cc.REF_BIN, tmp_path / "ref", seg_f, con_f, flags)
    n = cc.run_bin(cc.NEW_BIN, tmp_path / "new", seg_f, con_f, flags)
    assert n.returncode == 0, n.stderr[-2000:]
    assert r.returncode == 0, r.stderr[-2000:]
    only_ref, only_new, diff, n_files = cc.compare_dirs(tmp_path / "ref", tmp_path / "new")
    assert not only_ref and not only_new and not diff, (
        f"files only in reference: {only_ref[:5]}; only in ours: {only_new[:5]}; different: {diff[:5]}")
    return n_files


@pytest.mark.parametrize("name", sorted(cc.CASES))
def test_cli_matches_reference(tmp_path, name):
    _need_bins()
    kwargs, flags = cc.CASES[name]
    segs, contigs = cc.make_case(**kwargs)
    n_files = _run_both(tmp_path, segs, contigs, flags)
    if name.startswith("standard") or name.startswith("detection"):
        assert n_files > 2, "case produced no alignment files; it does not exercise the alignment phase"


def test_cli_contig_list(tmp_path):
    _need_bins()
    segs, contigs = cc.make_case(seed=11, n_contigs=80, n_related=16)
    lst = tmp_path / "list.txt"
    lst.write_text(" ".join(str(i) for i in range(1, 70)) + "\n")
    _run_both(tmp_path, segs, contigs, ["-nthreads=3", "-gen=2", f"-contig_list={lst}"])


def test_cli_fitting(tmp_path, monkeypatch):
    """OMPathFinder-style call (OMPathFinder.py:407-428): one compound segment map vs one sub-contig map, plus
    the KA-2 known answer of SURVEY.md section 8(c)."""
    _need_bins()
    monkeypatch.setenv("SA_SERVER", "0")  # in-process here; the default (resident server) has its own test below
    seg = np.array([0, 5200, 9100, 20500, 31000, 33900, 47000, 47001], dtype=np.float64)
    contig = np.array([0, 5150, 9300, 15000, 20400, 31200, 34000, 46800, 46801], dtype=np.float64)
    _run_both(tmp_path / "ka2", {1: seg}, {1: contig}, ["-fitting", "-min_labs=0", "-gen=2"])
    txt = (tmp_path / "ka2" / "new" / "SA_1_1_fitting_aln.txt").read_text().splitlines()
    assert txt[1] == "#1+\t51069\tFalse"
    rows = [l.split("\t") for l in txt[3:]]
    assert [(r[2], r[3], r[7]) for r in rows] == [("1", "1", "0"), ("2", "2", "9890.66"), ("3", "3", "19136.4"),
                                                   ("5", "4", "23197.7"), ("6", "5", "32258.9"),
                                                   ("7", "6", "42007.7"), ("8", "7", "51069")]
    rng = np.random.default_rng(3)
    for t in range(3):
        g = synth.genome_positions(rng, 500)
        s = synth.window_map(g, 10, int(rng.integers(5, 40)))
        c = synth.noisy_contig(rng, g, 8, int(rng.integers(10, 60)))
        _run_both(tmp_path / f"f{t}", {1: s, 2: synth.window_map(g, 100, 12)}, {1: c, 3: synth.noisy_contig(rng, g, 90, 30)},
                  ["-fitting", "-min_labs=0", "-gen=1"])


def test_cli_degenerate_shapes(tmp_path):
    """1-3 label contigs, 2-4 label segments (-min_labs=2), exact copies (zero discrepancies)."""
    _need_bins()
    segs, contigs = cc.make_edge_case(21)
    _run_both(tmp_path, segs, contigs, cc.EDGE_FLAGS)
    _run_both(tmp_path / "loc", segs, contigs, cc.EDGE_FLAGS + ["-local", "-no_tip_aln"])


def test_c1_gbm39_segments_vs_bspqi_assembly(tmp_path):
    """BASELINE.json configs[0] (/root/reference/README.md:162-191): the shipped GBM39 amplicon segment
    CMAP against a BspQI contig assembly -- 2,500 seeded synthetic contigs (seed 39) plus 6 noisy windows of the real
    hg19 chr7 map around EGFR -- standard mode, every output file byte-identical to the unmodified reference binary."""
    import os
    _need_bins()
    if not (cc.DATA / "hg19_BspQI.cmap.gz").exists():
        pytest.skip("oracle/_ref/data not present")
    seg_f, contigs = cc.make_c1()
    (tmp_path / "in").mkdir()
    synth.write_cmap(tmp_path / "in" / "contigs.cmap", contigs)
    flags = cc.C1_FLAGS + [f"-nthreads={min(os.cpu_count() or 1, 16)}"]
    r = cc.run_bin(cc.REF_BIN, tmp_path / "ref", seg_f, tmp_path / "in" / "contigs.cmap", flags)
    n = cc.run_bin(cc.NEW_BIN, tmp_path / "new", seg_f, tmp_path / "in" / "contigs.cmap", flags)
    assert r.returncode == 0 and n.returncode == 0, n.stderr[-2000:]
    only_ref, only_new, diff, n_files = cc.compare_dirs(tmp_path / "ref", tmp_path / "new")
    assert not only_ref and not only_new and not diff, (only_ref[:5], only_new[:5], diff[:5])
    assert n_files >= 8, "C1 must produce real alignments of the GBM39 segments on the chr7 contigs"


def test_c5_shaped_cohort_vs_reference(tmp_path, monkeypatch):
    """A C5-shaped subsample (BASELINE.json configs[4]: cohort sweep of many short segments x many short contigs) at a
    size where the cohort-scale host path takes its big-batch branches, byte-identical to the unmodified reference binary
    through BOTH alignment engines:
      * the device-resident loop (sa_align_rounds; default for standard runs): > 100,000 pairs in the tip pass, several
        replay workers, the packed path arena split into chunks (SA_ROUNDS_ARENA_BYTES);
      * the batched rounds (sa_align_batch_packed, what -local / -detection / oversized pairs use; forced here with
        SA_NO_ROUNDS=1): > 100,000 pairs per round (pinned pre-sizing, post-processing workers splicing private bitmaps
        into the arena), stored matrices split into many chunks (SA_STORE_BUDGET_CELLS)."""
    import os
    _need_bins()
    segs, contigs = cc.make_c5_shaped()
    seg_f, con_f = cc.write_case(tmp_path / "in", segs, contigs)
    flags = cc.C5_FLAGS + [f"-nthreads={min(os.cpu_count() or 1, 32)}"]
    r = cc.run_bin(cc.REF_BIN, tmp_path / "ref", seg_f, con_f, flags)
    assert r.returncode == 0, r.stderr[-2000:]
    monkeypatch.setenv("SA_VERBOSE", "1")
    # default settings first: pairs whose phase-1 score rules out an alignment file are skipped (exactly: see run_job)
    n = cc.run_bin(cc.NEW_BIN, tmp_path / "dflt", seg_f, con_f, flags + ["-ngpus=1"])
    assert n.returncode == 0, n.stderr[-2000:]
    only_ref, only_new, diff, n_files = cc.compare_dirs(tmp_path / "ref", tmp_path / "dflt")
    assert not only_ref and not only_new and not diff, (only_ref[:5], only_new[:5], diff[:5])
    skipped = [l for l in n.stderr.splitlines() if "can reach the score an alignment file needs" in l]
    assert len(skipped) == 2 and int(skipped[1].split()[0]) < int(skipped[1].split()[2]), skipped
    # ... then with every pair aligned (SA_NO_SKIP=1), which makes the tip pass big enough for the big-batch branches
    monkeypatch.setenv("SA_NO_SKIP", "1")
    monkeypatch.setenv("SA_ROUNDS_ARENA_BYTES", "20000000")
    n = cc.run_bin(cc.NEW_BIN, tmp_path / "new", seg_f, con_f, flags + ["-ngpus=1"])
    assert n.returncode == 0, n.stderr[-2000:]
    only_ref, only_new, diff, n_files = cc.compare_dirs(tmp_path / "ref", tmp_path / "new")
    assert not only_ref and not only_new and not diff, (only_ref[:5], only_new[:5], diff[:5])
    assert n_files > 300
    loops = [int(l.split()[3]) for l in n.stderr.splitlines() if l.strip().startswith("device-resident alignment loop:")]
    chunks = [l for l in n.stderr.splitlines() if "sa_align_rounds chunk" in l]
    assert len(loops) == 2 and max(loops) > 100000, f"device-resident loop sizes {loops}: the big-batch branches were not taken"
    assert len(chunks) > 3, "the path arena was not chunked"
    monkeypatch.delenv("SA_ROUNDS_ARENA_BYTES")
    monkeypatch.setenv("SA_NO_ROUNDS", "1")
    monkeypatch.setenv("SA_STORE_BUDGET_CELLS", "4000000")
    n = cc.run_bin(cc.NEW_BIN, tmp_path / "old", seg_f, con_f, flags + ["-ngpus=1"])
    assert n.returncode == 0, n.stderr[-2000:]
    only_ref, only_new, diff, n_files = cc.compare_dirs(tmp_path / "ref", tmp_path / "old")
    assert not only_ref and not only_new and not diff, (only_ref[:5], only_new[:5], diff[:5])
    rounds = [int(l.split()[2]) for l in n.stderr.splitlines() if l.strip().startswith("alignment round:")]
    chunks = [l for l in n.stderr.splitlines() if "sa_align_batch chunk" in l]
    assert max(rounds) > 100000, f"largest alignment round has {max(rounds)} pairs: the big-batch branches were not taken"
    assert len(chunks) > 2 * len(rounds), "stored matrices were not chunked"


@pytest.mark.parametrize("n_ctx", [2, 3])
def test_sharded_engine_matches_reference(tmp_path, monkeypatch, n_ctx):
    """The multi-GPU product path (Engine::shard / score / detect / align, host/main.cpp; reference fan-out
    SegAligner/main.cpp:531-551, 575-602): problems sharded by index over `n_ctx` contexts.  Runs on whatever the box has:
    with fewer than n_ctx GPUs the contexts share devices (SA_GPU_OVERSUBSCRIBE=1), which exercises exactly the same
    host code.  Standard and detection runs, byte-identical to the unmodified reference binary."""
    _need_bins()
    monkeypatch.setenv("SA_GPU_OVERSUBSCRIBE", "1")
    for name in ("standard_gen2", "detection"):
        kwargs, flags = cc.CASES[name]
        segs, contigs = cc.make_case(**kwargs)
        seg_f, con_f = cc.write_case(tmp_path / name / "in", segs, contigs)
        r = cc.run_bin(cc.REF_BIN, tmp_path / name / "ref", seg_f, con_f, flags)
        n = cc.run_bin(cc.NEW_BIN, tmp_path / name / "new", seg_f, con_f, flags + [f"-ngpus={n_ctx}"])
        assert r.returncode == 0 and n.returncode == 0, n.stderr[-2000:]
        only_ref, only_new, diff, n_files = cc.compare_dirs(tmp_path / name / "ref", tmp_path / name / "new")
        assert not only_ref and not only_new and not diff and n_files > 2, (only_ref[:5], only_new[:5], diff[:5])


def test_fitting_calls_are_fast_by_default(tmp_path, monkeypatch):
    """OMPathFinder.py:407-428 execs one `SegAligner -fitting` process per candidate path.  With NO environment set the
    first such call leaves the resident server behind (private socket directory, peer credentials checked) and the
    following calls are forwarded to it: 200 sequential calls must take < 1 s in total and write the same bytes as the
    unmodified reference binary."""
    import os
    import subprocess
    import time
    _need_bins()
    for k in ("SA_SERVER", "SA_SERVER_SOCKET", "SA_SERVER_IDLE"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("XDG_RUNTIME_DIR", str(tmp_path))  # keeps this test's server apart from any other
    os.chmod(tmp_path, 0o700)
    rng = np.random.default_rng(17)
    g = synth.genome_positions(rng, 2000)
    seg_f, con_f = cc.write_case(tmp_path / "in", {1: synth.window_map(g, 10, 9)}, {1: synth.noisy_contig(rng, g, 8, 14)})
    flags = ["-fitting", "-min_labs=0", "-gen=2"]
    ref = cc.run_bin(cc.REF_BIN, tmp_path / "ref", seg_f, con_f, flags)
    assert ref.returncode == 0
    try:
        warm = cc.run_bin(cc.NEW_BIN, tmp_path / "warm", seg_f, con_f, flags)  # starts the server (CUDA start-up)
        assert warm.returncode == 0, warm.stderr
        assert (tmp_path / "segaligner_b200" / "server.sock").exists()
        assert oct((tmp_path / "segaligner_b200").stat().st_mode & 0o777) == "0o700"
        outs = [tmp_path / f"o{k}" for k in range(200)]
        for o in outs:
            o.mkdir()
        t0 = time.perf_counter()
        for o in outs:
            rc = subprocess.call([str(cc.NEW_BIN), str(seg_f), str(con_f), f"-prefix={o}/SA"] + flags,
                                 stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            assert rc == 0
        dt = time.perf_counter() - t0
        want = cc.dir_digest(tmp_path / "ref")
        for o in outs:
            assert cc.dir_digest(o) == want
        assert dt < 1.0, f"200 fitting calls took {dt:.2f} s"
        print(f"200 sequential -fitting calls: {dt:.3f} s ({dt / 200 * 1e3:.2f} ms per call)")
    finally:
        subprocess.run([str(cc.NEW_BIN), "--server-stop"], capture_output=True, timeout=60)


def test_server_socket_is_not_trusted_blindly(tmp_path, monkeypatch):
    """A socket directory that is not a private directory of this user (here: group/world accessible) must not be used:
    the run then happens in-process and still succeeds with identical output."""
    import os
    _need_bins()
    for k in ("SA_SERVER_SOCKET", "SA_SERVER_IDLE"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("SA_SERVER", "1")
    monkeypatch.setenv("XDG_RUNTIME_DIR", str(tmp_path))
    os.chmod(tmp_path, 0o700)
    (tmp_path / "segaligner_b200").mkdir()
    os.chmod(tmp_path / "segaligner_b200", 0o777)  # somebody else could have planted a socket here
    seg = np.array([0, 5200, 9100, 20500, 31000, 33900, 47000, 47001], dtype=np.float64)
    contig = np.array([0, 5150, 9300, 15000, 20400, 31200, 34000, 46800, 46801], dtype=np.float64)
    seg_f, con_f = cc.write_case(tmp_path / "in", {1: seg}, {1: contig})
    p = cc.run_bin(cc.NEW_BIN, tmp_path / "new", seg_f, con_f, ["-fitting", "-min_labs=0", "-gen=2"])
    assert p.returncode == 0, p.stderr
    assert not (tmp_path / "segaligner_b200" / "server.sock").exists(), "a server was started in an untrusted directory"
    assert (tmp_path / "new" / "SA_1_1_fitting_aln.txt").read_text().splitlines()[1] == "#1+\t51069\tFalse"


def test_cli_resident_server(tmp_path):
    """SA_SERVER=1: the first call leaves a server holding the CUDA context on a Unix socket; later calls are
    forwarded to it.  Outputs (files and stdout) must equal the in-process run; repeated fitting calls -- the
    OMPathFinder pattern, one process per candidate path -- must be fast once the server is up."""
    import os
    import subprocess
    import time
    _need_bins()
    sock = str(tmp_path / "sa.sock")
    env = dict(os.environ, SA_SERVER="1", SA_SERVER_SOCKET=sock, SA_SERVER_IDLE="60")
    seg = np.array([0, 5200, 9100, 20500, 31000, 33900, 47000, 47001], dtype=np.float64)
    contig = np.array([0, 5150, 9300, 15000, 20400, 31200, 34000, 46800, 46801], dtype=np.float64)
    seg_f, con_f = cc.write_case(tmp_path / "in", {1: seg}, {1: contig})
    flags = ["-fitting", "-min_labs=0", "-gen=2"]
    direct = cc.run_bin(cc.NEW_BIN, tmp_path / "direct", seg_f, con_f, flags, env=dict(os.environ, SA_SERVER="0"))
    assert direct.returncode == 0
    try:
        times = []
        for k in range(6):
            (tmp_path / f"s{k}").mkdir()
            t0 = time.perf_counter()
            p = subprocess.run([str(cc.NEW_BIN), str(seg_f), str(con_f), f"-prefix={tmp_path}/s{k}/SA"] + flags,
                               capture_output=True, text=True, env=env, timeout=120)
            times.append(time.perf_counter() - t0)
            assert p.returncode == 0, p.stderr
            assert cc.dir_digest(tmp_path / f"s{k}") == cc.dir_digest(tmp_path / "direct")
            assert p.stdout.splitlines()[:3] == direct.stdout.splitlines()[:3]
        assert os.path.exists(sock)
        assert min(times[1:]) < 0.25, f"server round trips too slow: {times}"
        # a standard run through the server, relative paths resolved against the client's cwd
        segs, contigs = cc.make_case(seed=1)
        cc.write_case(tmp_path / "std", segs, contigs)
        (tmp_path / "std" / "a").mkdir()
        (tmp_path / "std" / "b").mkdir()
        f2 = ["-nthreads=3", "-min_labs=10", "-gen=2"]
        pa = subprocess.run([str(cc.NEW_BIN), "segs.cmap", "contigs.cmap", "-prefix=a/SA"] + f2, cwd=tmp_path / "std",
                            capture_output=True, text=True, env=env, timeout=300)
        pb = subprocess.run([str(cc.NEW_BIN), "segs.cmap", "contigs.cmap", "-prefix=b/SA"] + f2, cwd=tmp_path / "std",
                            capture_output=True, text=True, timeout=300)
        assert pa.returncode == 0 and pb.returncode == 0, pa.stderr + pb.stderr
        assert cc.dir_digest(tmp_path / "std" / "a") == cc.dir_digest(tmp_path / "std" / "b")
        bad = subprocess.run([str(cc.NEW_BIN), "segs.cmap", "contigs.cmap", "-gen=3"], cwd=tmp_path / "std",
                             capture_output=True, text=True, env=env, timeout=60)
        assert bad.returncode == 1 and "only valid arguments for -gen" in bad.stdout
    finally:
        subprocess.run([str(cc.NEW_BIN), "--server-stop"], env=env, capture_output=True, timeout=60)
    print("server round-trip times (s):", [round(t, 3) for t in times])
