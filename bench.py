#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native SegAligner DP hot path.

Metric (BASELINE.json): DP GCUPS (10^9 DP cells S[j][q] of the phase-1 scoring problems per second) and alignments/s,
whole job over N GPUs.  Workload at every N: BASELINE.json configs[1] = "synthetic batch: 2,000 graph segments x 500
Saphyr DLE1 contigs, both orientations" (seed 2000): ONE batch of ~1.9 M (segment+-, contig) problems.
With N ranks the batch is SHARDED BY PROBLEM INDEX (largest-first round-robin, dist_util.shard_problems; the reference's
fan-out is SegAligner/main.cpp:531-551), results are gathered to rank 0, and the line carries a checksum of the gathered
scores that must be identical for every N (strong scaling).

  value        phase-1 scoring kernels, device-timed (CUDA events on the library's stream, maps / problem list / results
               resident in HBM), max over ranks; all cells of the batch / that time.
  e2e          THE DROP-IN BOUNDARY: the executable bin/SegAligner (what ARAlignDetect.py / OMPathFinder.py exec) on the
               full workload from CMAP files to output files -- CUDA start-up, parse, phase-1 scoring, thresholds,
               alignment rounds, tip alignment, ~28 k *_aln.txt files -- on N GPUs (-ngpus=N: one process
               and one context per GPU, problems sharded by index).  Same definition as the reference arm and
               cpu_baseline: phase-1 scoring cells / process wall time.  Host<->device bytes are the library's own counters.
  whole_run_default_ngpus  (N > 1 lines) the same executable run without -ngpus: the executable's own choice of how many of
               the N visible GPUs a job of this size is worth (CUDA start-up of every extra process against its share of the
               work); e2e stays the forced -ngpus=N run.
  e2e_scoring  the scoring phase alone through the host-buffer C-ABI calls (sa_set_maps + sa_score_batch on this rank's
               shard, pinned host buffers, H2D/D2H inside) plus the gather of all results to rank 0.
  roofline     the binding resource of the executed work, the XU (MUFU) pipe: 2 MUFU per predecessor candidate (fp32 score
               bound) + 2 conversions per bit-exact replay, against the MUFU issue peak measured live on this GPU.  The
               algorithmic FP64-issue figure SURVEY 8(d) defines (18 fp64 instr per predecessor evaluation, un-pruned) is
               reported beside it as roofline.fp64_algorithmic (it exceeds 1: the kernel prunes exactly).
  cpu_baseline the UNMODIFIED reference SegAligner (oracle/_ref/SegAligner_ref) on a seeded RANDOM sample of the same
               workload on this box's host cores: whole run, and -no_tip_aln (scoring-dominated) for a phase-matched ratio.
  c3_detection (N=1 line) the north-star target workload: hg19 BspQI detection sweep (64 synthetic regions x 2
               orientations vs the whole reference genome CMAP, `-detection`), whole process, with the reference timed on
               a byte-compared subset.

`--impl reference` times only the reference binary (rank 0), same metric / unit / config.workload.
"""
from __future__ import annotations

import argparse
import gzip
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# The driver reads ONE JSON line from stdout.  Libraries loaded later (NCCL prints its version banner to stdout on some
# boxes) must not be able to add lines: keep a private duplicate of the original stdout for the JSON line and point
# file descriptor 1 at stderr for everything else.
_JSON_FD = os.dup(1)
os.dup2(2, 1)


def emit_json(obj) -> None:
    os.write(_JSON_FD, (json.dumps(obj) + "\n").encode())


def log(msg: str) -> None:
    sys.stderr.write(f"[bench] {msg}\n")
    sys.stderr.flush()


METRIC = "dp_gcups"
UNIT = "GCUPS"
REF_BIN = ROOT / "oracle" / "_ref" / "SegAligner_ref"
NEW_BIN = ROOT / "ampliconreconstructorom_b200" / "bin" / "SegAligner"
DATA = ROOT / "oracle" / "_ref" / "data"
N_SEGS, N_CONTIGS, SEED = 2000, 500, 2000
WORKLOAD = "C2 synthetic batch: 2,000 graph segments x 500 Saphyr DLE1 contigs, both orientations (SURVEY 8(d), seed 2000)"
C2_FLAGS = ["-min_labs=12", "-gen=2"]
DRAM_BYTES_PER_LAUNCH = 13.5e6  # ncu dram__bytes_read.sum + dram__bytes_write.sum of the scoring kernel (profiles/r1_fill_kernel_pruned_ncu.txt)


def pair_evals(n_labels: np.ndarray) -> np.ndarray:
    """sum_{j<n} min(j, 6) per map: the look-back window sizes of dp_align (SURVEY 8(d) work model)."""
    n = n_labels.astype(np.int64)
    full = np.maximum(n - 6, 0)
    head = np.minimum(n, 6)
    return full * 6 + head * (head - 1) // 2


def c2_maps(scale: float):
    """The C2 workload (the same on every rank): ({seg id: map}, {contig id: map}).  Pure numpy (synth.py)."""
    from ampliconreconstructorom_b200 import synth
    n_segs = max(int(round(N_SEGS * scale)), 4)
    n_contigs = max(int(round(N_CONTIGS * scale)), 40)
    return synth.make_workload(SEED, n_segs, n_contigs)


def scoring_cells(segs: dict, contigs: dict, min_labs: int = 12) -> float:
    """phase-1 DP cells of a SegAligner run: both orientations of every segment make_reverse_cmap keeps x every contig"""
    nx = sum(len(m) - 1 for m in segs.values() if len(m) - 1 >= min_labs)
    nb = sum(len(m) - 1 for m in contigs.values())
    return 2.0 * nx * nb


def write_inputs(d: Path, segs: dict, contigs: dict):
    from ampliconreconstructorom_b200 import synth
    d.mkdir(parents=True, exist_ok=True)
    synth.write_cmap(d / "segs.cmap", segs)
    synth.write_cmap(d / "contigs.cmap", contigs)
    return d / "segs.cmap", d / "contigs.cmap"


def clock_sampler(gpu_index: int, stop: threading.Event, out: list):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                              "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except OSError:
        return
    while not stop.is_set():
        line = p.stdout.readline()
        if not line:
            break
        f = [x.strip() for x in line.split(",")]
        if len(f) >= 8:
            out.append(f)
    p.terminate()


def summarise_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    sm = [float(s[1]) for s in samples if s[1].replace(".", "").isdigit()]
    mx = [float(s[2]) for s in samples if s[2].replace(".", "").isdigit()]
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for k, n in enumerate(names) if any(s[4 + k].lower().startswith("active") for s in samples)]
    pw = [float(s[3]) for s in samples if s[3].replace(".", "").isdigit()]
    return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(samples)}


# ---- the reference binary on a bounded random sample -----------------------------------------------------------------
def run_binary(binary: Path, seg_f, con_f, prefix: str, flags, env=None, capture=False):
    """One process, wall-clock timed from the outside (the program's own `elapsed seconds` is CPU time)."""
    t0 = time.perf_counter()
    p = subprocess.run([str(binary), str(seg_f), str(con_f), f"-prefix={prefix}"] + list(flags), stdout=subprocess.DEVNULL,
                       stderr=subprocess.PIPE if capture else subprocess.DEVNULL, text=True, env=env)
    dt = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError(f"{binary.name} exited {p.returncode}: {(p.stderr or '')[-400:]}")
    return dt, (p.stderr or "")


def reference_sample(segs: dict, contigs: dict, target_s: float, n_threads: int, seed: int = 12345):
    """A seeded RANDOM subset of the workload's segments and contigs sized so that one whole reference run needs about
    target_s seconds on this box (probe run first; the reference's cost is linear in scoring cells)."""
    rng = np.random.default_rng(seed)
    seg_ids = [int(i) for i in rng.permutation(sorted(segs))]
    con_ids = [int(i) for i in rng.permutation(sorted(contigs))]
    # >= 4 contigs per reference thread (it splits the contig set over its threads), >= 64 so thresholds are defined
    n_con = min(len(con_ids), max(64, 4 * n_threads))
    sub_c = {i: contigs[i] for i in con_ids[:n_con]}
    probe_s = {i: segs[i] for i in [s for s in seg_ids if len(segs[s]) - 1 >= 12][:3]}
    with tempfile.TemporaryDirectory() as d:
        sf, cf = write_inputs(Path(d), probe_s, sub_c)
        dt, _ = run_binary(REF_BIN, sf, cf, f"{d}/SA", C2_FLAGS + [f"-nthreads={n_threads}"])
    rate = scoring_cells(probe_s, sub_c) / max(dt, 1e-3)
    # the tip phase re-aligns every segment against every contig that received an alignment, so the whole-run cost per
    # scoring cell grows with the number of segments in the sample (measured: 2.4x from the 3-segment probe to ~50 segments)
    want = rate * target_s * 0.4
    sub_s = {}
    for i in seg_ids:
        sub_s[i] = segs[i]
        if scoring_cells(sub_s, sub_c) >= want:
            break
    return sub_s, sub_c


def time_reference(segs: dict, contigs: dict, n_threads: int, steps: int = 1, with_no_tip: bool = True):
    """Whole-run wall times (and the -no_tip_aln, scoring-dominated run) of the unmodified reference on the sample."""
    whole, no_tip = [], []
    with tempfile.TemporaryDirectory() as d:
        sf, cf = write_inputs(Path(d), segs, contigs)
        for it in range(steps):
            dt, _ = run_binary(REF_BIN, sf, cf, f"{d}/w{it}_SA", C2_FLAGS + [f"-nthreads={n_threads}"])
            whole.append(dt)
            if with_no_tip:
                dt, _ = run_binary(REF_BIN, sf, cf, f"{d}/n{it}_SA", C2_FLAGS + [f"-nthreads={n_threads}", "-no_tip_aln"])
                no_tip.append(dt)
    return whole, no_tip


def cpu_baseline_block(segs_all, contigs_all, target_s, steps=1, with_b200=False):
    n_threads = os.cpu_count() or 1
    total = scoring_cells(segs_all, contigs_all)
    segs, contigs = reference_sample(segs_all, contigs_all, target_s, n_threads)
    cells = scoring_cells(segs, contigs)
    whole, no_tip = time_reference(segs, contigs, n_threads, steps)
    tw, tn = sum(whole) / len(whole), sum(no_tip) / len(no_tip)
    n_prob = 2 * sum(1 for m in segs.values() if len(m) - 1 >= 12) * len(contigs)
    sample = (f"unmodified reference SegAligner (oracle/_ref), seeded random sample of {len(segs)} of {len(segs_all)} segments x "
              f"{len(contigs)} of {len(contigs_all)} contigs = {cells:.3e} scoring cells ({100 * cells / total:.2f} % of the "
              f"workload), -nthreads={n_threads}; whole run (parse + scoring + thresholds + alignment + tip alignment + files) "
              f"{tw:.1f} s wall; with -no_tip_aln {tn:.1f} s")
    block = {"value": cells / tw / 1e9, "unit": UNIT, "cores": n_threads, "kind": "reference", "sample": sample,
             "sample_fraction_of_cells": cells / total, "wall_s": tw,
             "scoring_dominated": {"value": cells / tn / 1e9, "unit": UNIT, "wall_s": tn,
                                   "what": "same sample with -no_tip_aln: parse + phase-1 scoring + thresholds + alignment of the "
                                           "passing pairs; phase-matched to e2e_scoring"}}
    if with_b200 and NEW_BIN.exists():
        # the drop-in executable on the IDENTICAL sample files (start-up dominated at this size, but exactly like for like)
        with tempfile.TemporaryDirectory() as d:
            sf, cf = write_inputs(Path(d), segs, contigs)
            best = None
            for rep in range(2):
                dt, _ = run_binary(NEW_BIN, sf, cf, f"{d}/b{rep}_SA", C2_FLAGS + [f"-nthreads={n_threads}", "-ngpus=1"])
                best = dt if best is None else min(best, dt)
        block["b200_on_same_sample"] = {"value": cells / best / 1e9, "unit": UNIT, "wall_s": best, "speedup": tw / best,
                                        "what": "bin/SegAligner whole process on the same sample files, 1 GPU (CUDA start-up included)"}
    block["note"] = ("the sample under-states the reference's cost on the full workload: its tip phase re-aligns every segment "
                     "against every contig that received an alignment, which is nearly all contigs for 2,000 segments but few for a "
                     "small sample (whole/no-tip = %.2f here vs ~4.5 on the full batch)" % (tw / tn))
    return block, whole, cells, n_prob


def run_reference_arm(args, rank):
    if rank != 0:
        return
    if not REF_BIN.exists():
        emit_json({"impl": "reference", "unavailable": "oracle/_ref/SegAligner_ref was not built"})
        return
    segs_all, contigs_all = c2_maps(args.scale)
    steps = max(1, args.steps)
    per_step = max(20.0, min(args.ref_seconds, args.ref_total_seconds / steps))
    block, whole, cells, n_prob = cpu_baseline_block(segs_all, contigs_all, per_step, steps)
    t = sum(whole)
    gcups = cells * len(whole) / t / 1e9
    block["value"] = gcups
    emit_json({"impl": "reference", "metric": METRIC, "value": gcups, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
               "warmup": args.warmup, "warmup_executed": 0, "ms_per_step": 1e3 * t / len(whole), "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f32 scores, f64 inside libm powf", "data": "synthetic",
               "config": {"workload": WORKLOAD, "scale": args.scale,
                          "what": "whole SegAligner process, wall clock; value = phase-1 scoring cells / wall (same definition as "
                                  "the B200 arm's e2e); CPU runs need no warm-up steps",
                          "sample": block["sample"]},
               "alignments_per_s": n_prob * len(whole) / t,
               "cpu_baseline": block,
               "e2e": {"value": gcups, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gpu_launches": 0})


# ---- the drop-in executable, whole process -----------------------------------------------------------------------------
def parse_verbose(err: str):
    """SA_VERBOSE=1 stderr of bin/SegAligner -> (phase seconds by name, H2D bytes, D2H bytes, GPUs the run used or None)."""
    phases, h2d, d2h, used = {}, 0, 0, None
    for line in err.splitlines():
        if line.startswith("SegAligner(b200): copied H2D"):
            t = line.split()
            h2d, d2h = int(t[3]), int(t[6])
        elif line.startswith("SegAligner(b200): wall") and line.rstrip().endswith("GPU(s)"):
            try:
                used = int(line.split()[-2])
            except ValueError:
                pass
        elif line.startswith("SegAligner(b200):") and line.rstrip().endswith(" s"):
            name = line[len("SegAligner(b200):"):].rsplit(None, 2)[0].strip()
            try:
                phases[name] = float(line.split()[-2])
            except ValueError:
                pass
    return phases, h2d, d2h, used


def time_executable(segs, contigs, n_gpus: int, reps: int = 1):
    """bin/SegAligner on the given maps, whole process (every phase, every output file); returns the JSON block.
    n_gpus = 0: no -ngpus flag, the executable picks the number of GPUs from the job size (its default behaviour)."""
    cells = scoring_cells(segs, contigs)
    ng_flag = [f"-ngpus={n_gpus}"] if n_gpus > 0 else []
    env = dict(os.environ, SA_VERBOSE="1")
    with tempfile.TemporaryDirectory() as d:
        d = Path(d)
        sf, cf = write_inputs(d, segs, contigs)
        # warm the binary, the CUDA driver and the file system on a tiny subset (not timed)
        wsegs = {i: segs[i] for i in sorted(segs)[:4]}
        wcont = {i: contigs[i] for i in sorted(contigs)[:40]}
        wf = write_inputs(d / "warm", wsegs, wcont)
        run_binary(NEW_BIN, wf[0], wf[1], f"{d}/warm/SA", C2_FLAGS + ["-nthreads=8"] + ng_flag, env=env, capture=True)
        best, err = None, ""
        for rep in range(reps):
            (d / f"r{rep}").mkdir()
            dt, e = run_binary(NEW_BIN, sf, cf, f"{d}/r{rep}/SA", C2_FLAGS + [f"-nthreads={os.cpu_count() or 1}"] + ng_flag,
                               env=env, capture=True)
            if best is None or dt < best:
                best, err = dt, e
        n_files = len([f for f in (d / "r0").iterdir() if f.name.endswith("_aln.txt")])
    phases, h2d, d2h, used = parse_verbose(err)
    return {"value": cells / best / 1e9, "unit": UNIT, "wall_s": best, "scoring_cells": cells, "n_gpus": n_gpus if n_gpus > 0 else used,
            "alignment_files": n_files, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "phases_s": phases,
            "what": "bin/SegAligner whole process from CMAP files to output files (CUDA start-up, parse, phase-1 scoring, thresholds, "
                    "alignment rounds, tip alignment, all *_aln.txt files); value = phase-1 scoring cells / process wall time"}


def c3_detection(n_regions=64, n_ref_regions=2):
    """North-star target workload (the ARAlignDetect.py:408-411 call): 64 synthetic unaligned regions
    (21-500 labels, seed 3: windows of the real hg19 BspQI map through the optical-map error model) searched against the whole
    hg19 BspQI reference CMAP with -detection -gen=1.  Whole process of the drop-in on the full sweep; the unmodified
    reference on a seeded subset whose output files are byte-compared with the drop-in's on the same subset."""
    from ampliconreconstructorom_b200 import capi, synth
    if not (DATA / "hg19_BspQI.cmap.gz").exists():
        return {"unavailable": "oracle/_ref/data/hg19_BspQI.cmap.gz not present"}
    import hashlib
    tmp = Path(tempfile.mkdtemp(prefix="c3_"))
    try:
        hg = tmp / "hg19.cmap"
        with gzip.open(DATA / "hg19_BspQI.cmap.gz", "rb") as f, open(hg, "wb") as g:
            shutil.copyfileobj(f, g)
        ids, off, pos = capi.host_parse_cmap(hg)
        labels = np.diff(off) - 1
        rng = np.random.default_rng(3)
        big = [k for k in range(len(ids)) if labels[k] > 2000]
        regions = {}
        for r in range(n_regions):
            k = big[int(rng.integers(0, len(big)))]
            n = int(rng.integers(21, 501))
            chrom = pos[off[k]:off[k + 1] - 1].astype(np.float64)
            s = int(rng.integers(1, chrom.size - n - 2))
            regions[r + 1] = synth.noisy_contig(rng, chrom, s, n)
        total_nx = sum(len(m) - 1 for m in regions.values())
        cells = 2.0 * total_nx * float(labels.sum())
        synth.write_cmap(tmp / "regions.cmap", regions)
        flags = [f"-nthreads={os.cpu_count() or 1}", "-min_labs=10", "-gen=1", "-detection"]
        env = dict(os.environ, SA_VERBOSE="1")
        best = None
        for rep in range(2):  # the first run warms the driver and the page cache
            (tmp / f"o{rep}").mkdir()
            dt, err = run_binary(NEW_BIN, tmp / "regions.cmap", hg, f"{tmp}/o{rep}/SA", flags + ["-ngpus=1"], env=env, capture=True)
            best = dt if best is None else min(best, dt)
        phases = {}
        for line in err.splitlines():
            if line.startswith("SegAligner(b200):") and line.rstrip().endswith(" s"):
                try:
                    phases[line[len("SegAligner(b200):"):].rsplit(None, 2)[0].strip()] = float(line.split()[-2])
                except ValueError:
                    pass
        n_aln = len([f for f in (tmp / "o1").iterdir() if f.name.endswith("_aln.txt")])
        out = {"workload": f"hg19 BspQI detection sweep: {n_regions} regions ({total_nx} labels) x 2 orientations vs hg19_BspQI.cmap "
                           f"({int(labels.sum())} labels), -detection -gen=1",
               "cells": cells, "wall_s": best, "value": cells / best / 1e9, "unit": UNIT, "alignment_files": n_aln, "phases_s": phases,
               "what": "bin/SegAligner whole process, 1 GPU"}
        if REF_BIN.exists():
            pick = sorted(int(i) for i in rng.permutation(sorted(regions))[:n_ref_regions])
            sub = {k: regions[k] for k in pick}
            synth.write_cmap(tmp / "sub.cmap", sub)
            sub_cells = 2.0 * sum(len(m) - 1 for m in sub.values()) * float(labels.sum())
            (tmp / "r").mkdir()
            (tmp / "n").mkdir()
            dtr, _ = run_binary(REF_BIN, tmp / "sub.cmap", hg, f"{tmp}/r/SA", flags)
            run_binary(NEW_BIN, tmp / "sub.cmap", hg, f"{tmp}/n/SA", flags + ["-ngpus=1"])
            da = {f.name: hashlib.sha256(f.read_bytes()).hexdigest() for f in (tmp / "r").iterdir()}
            db = {f.name: hashlib.sha256(f.read_bytes()).hexdigest() for f in (tmp / "n").iterdir()}
            out["reference"] = {"value": sub_cells / dtr / 1e9, "unit": UNIT, "cores": os.cpu_count() or 1, "wall_s": dtr,
                                "sample": f"unmodified reference on {len(sub)} of the {n_regions} regions (seeded pick {pick}), "
                                          f"{sub_cells:.3e} cells ({100 * sub_cells / cells:.1f} % of the sweep)",
                                "outputs_byte_identical_to_b200_on_sample": da == db, "files_compared": len(da)}
            out["speedup_whole_process"] = (cells / best) / (sub_cells / dtr)
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=float(os.environ.get("SA_BENCH_SCALE", "1.0")),
                    help="linear scale of both map counts (1.0 = the named C2 config; <1 only for dry runs)")
    ap.add_argument("--ref-seconds", type=float, default=90.0, help="wall seconds of one whole reference run (sample size)")
    ap.add_argument("--ref-total-seconds", type=float, default=240.0, help="cap on the reference arm's whole-run seconds over all steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip whole-executable e2e, cpu_baseline and c3_detection (dry runs)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from ampliconreconstructorom_b200 import capi, dist_util

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    gloo = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        gloo = dist.new_group(backend="gloo")  # host-side gather of the result records (no device collective is needed)
    ctx = capi.Context(local_rank)

    # ---- the one batch (identical on every rank) and this rank's shard of it
    segs_all, contigs_all = c2_maps(args.scale)
    sid, soff, spos = capi.flatten_maps(segs_all)
    fid, foff, fpos = capi.host_make_reverse(sid, soff, spos, 12)  # both orientations (make_reverse_cmap)
    cid, coff, cpos = capi.flatten_maps(contigs_all)
    fprob = capi.host_non_collapse_probs(foff, fpos, 2)
    cprob = capi.host_non_collapse_probs(coff, cpos, 2)

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()

    keep, bufs = [], {}
    for k, v in (("foff", foff), ("fpos", fpos), ("fprob", fprob), ("coff", coff), ("cpos", cpos), ("cprob", cprob)):
        t, a = pinned(v)
        keep.append(t)
        bufs[k] = a
    n_c, n_s = len(cid), len(fid)
    nb = (np.diff(coff) - 1).astype(np.int64)
    nx = (np.diff(foff) - 1).astype(np.int64)
    cc, ss = np.meshgrid(np.arange(n_c, dtype=np.int32), np.arange(n_s, dtype=np.int32), indexing="ij")
    cc, ss = cc.ravel(), ss.ravel()
    n_all = cc.size
    cells_all = float(nb.sum()) * float(nx.sum())
    mine = np.sort(dist_util.shard_problems(nb[cc] * nx[ss], world, rank))  # indices into the global problem list
    tp, problems = pinned(ctx.problems(cc[mine], ss[mine]).view(np.uint8))
    keep.append(tp)
    problems = problems.view(capi.problem_dtype)
    n_prob = problems.size
    cells = float((nb[cc[mine]] * nx[ss[mine]]).sum())
    evals = float((pair_evals(nb)[cc[mine]] * pair_evals(nx)[ss[mine]]).sum())
    mode = ctx.mode(capi.SA_INIT_SEMIGLOBAL, capi.SA_START_SEMIGLOBAL, 0, 6, 0)
    log(f"rank {rank}/{world}: {n_prob} of {n_all} problems, {cells:.3e} of {cells_all:.3e} cells")

    # ---- resident inputs for the device-timed number
    ctx.set_maps(capi.SA_SIDE_CONTIG, bufs["coff"], bufs["cpos"], bufs["cprob"])
    ctx.set_maps(capi.SA_SIDE_SEG, bufs["foff"], bufs["fpos"], bufs["fprob"])
    plan = ctx.plan_create(problems)
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fp64_peak = ctx.fp64_peak()
    mufu_peak = ctx.mufu_peak()

    def barrier():
        torch.cuda.synchronize()
        ctx.sync()
        if world > 1:
            dist.barrier()

    def step():
        l2_flush.zero_()          # flush L2 between iterations (torch stream)
        torch.cuda.synchronize()  # ... completed before the library's stream starts the step
        plan.score(mode)

    for _ in range(args.warmup):
        step()
    barrier()
    samples, stop = [], threading.Event()
    th = threading.Thread(target=clock_sampler, args=(local_rank, stop, samples), daemon=True)
    th.start()
    fill_ms, launches = [], 0
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        l2_flush.zero_()
        torch.cuda.synchronize()
        ctx.event_record(0)
        plan.score(mode)
        ctx.event_record(1)
        dev_ms += ctx.event_elapsed_ms()
        tm = ctx.timing()
        fill_ms.append(tm.fill_ms)
        launches += tm.n_launches
    barrier()
    wall = time.perf_counter() - t0
    stop.set()
    res = plan.results()
    info = plan.info()
    n_replays, n_rescans = ctx.prune_stats()  # work the last launch actually executed (exact pruning)

    # ---- scoring phase end to end through the host-buffer C-ABI calls + gather of every rank's results to rank 0
    def gather_results(local):
        """rank 0 returns the result records of the whole batch in problem order; other ranks return None"""
        if world == 1:
            full = np.empty(n_all, dtype=capi.result_dtype)
            full[mine] = local
            return full
        cap = (n_all + world - 1) // world
        pad = np.zeros(cap, dtype=capi.result_dtype)
        pad[:local.size] = local
        send = torch.from_numpy(pad.view(np.uint8).copy())
        recv = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
        dist.gather(send, recv, dst=0, group=gloo)
        if rank != 0:
            return None
        full = np.empty(n_all, dtype=capi.result_dtype)
        cells_p = nb[cc] * nx[ss]
        for r in range(world):
            idx = np.sort(dist_util.shard_problems(cells_p, world, r))
            full[idx] = recv[r].numpy().view(capi.result_dtype)[:idx.size]
        return full

    t_e2e = 0.0
    e2e_steps = max(1, min(args.steps, 2))
    h2d0, d2h0 = ctx.traffic()
    full = None
    for it in range(e2e_steps + 1):  # first pass untimed
        barrier()
        if it == 1:
            h2d0, d2h0 = ctx.traffic()
        t1 = time.perf_counter()
        ctx.set_maps(capi.SA_SIDE_CONTIG, bufs["coff"], bufs["cpos"], bufs["cprob"])
        ctx.set_maps(capi.SA_SIDE_SEG, bufs["foff"], bufs["fpos"], bufs["fprob"])
        res2 = ctx.score_batch(mode, problems)
        full = gather_results(res2)
        if world > 1:
            dist.barrier(group=gloo)
        if it > 0:
            t_e2e += time.perf_counter() - t1
    h2d1, d2h1 = ctx.traffic()
    assert np.array_equal(res2["score"].view(np.uint32), res["score"].view(np.uint32)), "e2e and resident results differ"
    checksum = None
    if rank == 0:
        sc = full["score"]
        checksum = {"sum_scores": float(np.sum(sc.astype(np.float64))),
                    "xor_score_bits": int(np.bitwise_xor.reduce(sc.view(np.uint32))),
                    "sum_start_cells": int(full["j"].astype(np.int64).sum() + full["q"].astype(np.int64).sum()),
                    "n": int(full.size)}

    # ---- reductions over ranks (max time, summed work)
    (dev_ms_max, wall_ms_max, e2e_ms_max, fill_sum_max), (cells_sum, prob_sum, evals_sum, launches_sum, h2d_sum, d2h_sum,
                                                         replay_sum, rescan_sum) = \
        dist_util.reduce_over_ranks([dev_ms, wall * 1e3, t_e2e * 1e3 / e2e_steps, sum(fill_ms)],
                                    [cells, float(n_prob), evals, float(launches), float(h2d1 - h2d0) / e2e_steps,
                                     float(d2h1 - d2h0) / e2e_steps, float(n_replays), float(n_rescans)], device="cuda")
    plan.destroy()
    ctx.close()
    del l2_flush
    torch.cuda.empty_cache()

    # ---- the drop-in boundary: the executable on the whole workload, rank 0 drives all N GPUs (-ngpus=N)
    exe = exe_auto = None
    if not args.no_extras:
        if world > 1:
            dist.barrier(group=gloo)  # host-side wait: an NCCL barrier would keep a kernel spinning on the idle ranks' GPUs
        if rank == 0:
            log(f"whole executable on the full workload, -ngpus={world} ...")
            try:
                exe = time_executable(segs_all, contigs_all, world, reps=1 if args.steps < 3 else 2)
            except RuntimeError as e:
                exe = {"value": None, "unit": UNIT, "error": str(e)[-400:]}
            if world > 1:
                # beside the forced -ngpus=N run: what the executable does when nobody tells it how many GPUs to use (it weighs
                # the CUDA start-up of every additional process against its share of the work)
                try:
                    log("whole executable on the full workload, no -ngpus flag (the executable's own choice) ...")
                    exe_auto = time_executable(segs_all, contigs_all, 0, reps=1)
                    exe_auto["what"] = "the same run without -ngpus: the executable picks the number of GPUs from the job size"
                except Exception as e:  # noqa: BLE001 -- an optional key must not cost the line
                    exe_auto = {"value": None, "unit": UNIT, "error": str(e)[-400:]}
        if world > 1:
            dist.barrier(group=gloo)

    if rank == 0:
        K = args.steps
        clocks = summarise_clocks(samples)
        sm_hz = 1e6 * max(float(clocks.get("sm_mhz") or 1965.0), 1.0)
        value = cells_sum * K / (dev_ms_max * 1e-3) / 1e9
        t_kernel = fill_sum_max * 1e-3 / K                       # slowest rank's kernel time per launch
        xu_instr = 2.0 * evals_sum + 2.0 * replay_sum             # whole job per step: MUFU of the bound pass + F2F of the replays
        xu_rate = xu_instr / world / t_kernel                     # per GPU
        fp64_alg = 18.0 * evals_sum / world / t_kernel
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 scores, f64 inside powf (bit-exact glibc replay)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "scale": args.scale,
                       "phase_timed_by_value": "phase-1 scoring: semi-global init, lookback-6 dp_align, backtrack-start scan",
                       "segments_both_orientations": int(n_s), "contigs": int(n_c), "problems": int(n_all), "cells": cells_all,
                       "sharding": "ONE batch sharded by problem index over the ranks (largest-first round-robin), results gathered to "
                                   "rank 0; no device collective (problems are independent)",
                       "l2": "flushed between steps (256 MiB write)", "large_problems_cooperative": info["n_large"]},
            "alignments_per_s": prob_sum * K / (dev_ms_max * 1e-3),
            "wall_ms_per_step": wall_ms_max / K,
            "gpu_launches": int(launches_sum),
            "checksum": checksum,
            "e2e_scoring": {"value": cells_sum / (e2e_ms_max * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d_sum),
                            "d2h_bytes_per_step": int(d2h_sum), "alignments_per_s": prob_sum / (e2e_ms_max * 1e-3), "steps": e2e_steps,
                            "what": "sa_set_maps + sa_score_batch on each rank's shard (pinned host buffers in, host results out) + "
                                    "gather of all result records to rank 0; bytes = the library's own copy counters, summed over ranks"},
            "roofline": {"bound": "xu_pipe (MUFU issue)", "achieved": xu_rate / 1e12, "peak": mufu_peak / 1e12,
                         "unit": "T XU thread-instr/s per GPU", "frac": xu_rate / mufu_peak,
                         "peak_source": "live MUFU lg2/ex2 microbenchmark on this GPU (sa_measure_mufu_peak); MEASURED_PEAKS.json has "
                                        "neither an fp64 nor an XU entry",
                         "work": "2 MUFU (lg2, ex2) per predecessor candidate for the fp32 score bound + 2 conversions per bit-exact "
                                 "replay (DESIGN.md 4.1); executed counts: candidates from the work model, replays from the kernel's counter",
                         "kernel_ms_per_launch": t_kernel * 1e3,
                         "traffic": DRAM_BYTES_PER_LAUNCH,
                         "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per scoring launch on 1 GPU "
                                           "(profiles/); HBM is not a bound: results are 16 B per problem",
                         "fp64_algorithmic": {"achieved": fp64_alg / 1e12, "peak": fp64_peak / 1e12,
                                              "unit": "T fp64-instr/s per GPU (thread-level DFMA/DADD/DMUL)",
                                              "speedup_vs_dense_fp64_bound": fp64_alg / fp64_peak,
                                              "note": "SURVEY 8(d) algorithmic work: 18 fp64 instr x every predecessor evaluation of the "
                                                      "reference (un-pruned) / kernel time / live DFMA peak.  > 1 because the kernel prunes "
                                                      "exactly (fp32 bounds for all candidates, one bit-exact replay per cell)"},
                         "executed": {"powf_replays_per_step": replay_sum, "cells_rescanned_per_step": rescan_sum,
                                      "replays_per_cell": replay_sum / cells_sum,
                                      "fp64_frac_of_peak": 17.0 * replay_sum / world / t_kernel / fp64_peak,
                                      "xu_frac_of_nominal_16_lanes": xu_rate / (16.0 * 148 * sm_hz)}},
            "clocks": clocks,
        }
        if exe is not None:
            out["e2e"] = {"value": exe.get("value"), "unit": UNIT, "h2d_bytes_per_step": exe.get("h2d_bytes_per_step", 0),
                          "d2h_bytes_per_step": exe.get("d2h_bytes_per_step", 0), "steps": 1, **{k: v for k, v in exe.items()
                                                                                                   if k not in ("value", "unit")}}
            out["whole_run" if world == 1 else "whole_run_ngpus"] = exe
            if exe_auto is not None:
                out["whole_run_default_ngpus"] = exe_auto
        else:
            out["e2e"] = dict(out["e2e_scoring"], note="--no-extras: the whole-executable run was skipped; this is e2e_scoring")
        if world == 1 and not args.no_extras:
            if args.no_cpu_baseline or not REF_BIN.exists():
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                       "sample": "skipped (--no-cpu-baseline)" if args.no_cpu_baseline else "oracle/_ref/SegAligner_ref not present"}
            else:
                log("cpu_baseline: reference SegAligner on a random sample ...")
                out["cpu_baseline"] = cpu_baseline_block(segs_all, contigs_all, args.ref_seconds, 1, with_b200=True)[0]
            log("c3_detection ...")
            try:
                out["c3_detection"] = c3_detection()
            except RuntimeError as e:
                out["c3_detection"] = {"error": str(e)[-400:]}
        emit_json(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
