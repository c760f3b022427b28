#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native SegAligner DP hot path.

Metric (BASELINE.json): DP GCUPS (10^9 DP cells S[j][q] per second) and alignments/s, whole job over N GPUs.
Workload at every N: BASELINE.json configs[1] = "synthetic batch: 2,000 graph segments x 500 Saphyr DLE1
contigs, both orientations" (seed 2000+rank): 2,000,000 phase-1 scoring problems per GPU
(semi-global init, banded dp_align, backtrack-start scan), weak scaling: every rank scores its own C2 batch.

  value   device-timed: maps, problem list and results resident in HBM (sa_plan_score only), CUDA events on the
          library's stream, max over ranks.
  e2e     the same work through the host-buffer C-ABI calls a caller makes (sa_set_maps + sa_score_batch):
          pinned host buffers, H2D of maps/problems and D2H of results inside the timed region.
  roofline  FP64-issue bound: algorithmic 18 fp64 instructions per predecessor evaluation (glibc powf replay)
          / fill-kernel time / DFMA issue peak measured live on this GPU (MEASURED_PEAKS.json has no fp64 entry).
  cpu_baseline  the UNMODIFIED reference SegAligner (oracle/_ref/SegAligner_ref) on a bounded sample of the same
          workload on the host cores of this box, -nthreads=all.

  whole_run  the drop-in executable (bin/SegAligner) on a 30 % x 30 % subsample, whole process with every phase and all
          output files -- the same definition as cpu_baseline (phase-1 scoring cells / process wall time), so that
          whole_run / cpu_baseline compares like with like; `value` and `e2e` are the scoring phase alone.

`--impl reference` times only that reference binary (rank 0), same metric/unit/config.
"""
from __future__ import annotations

import argparse
import json
import os
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


METRIC = "dp_gcups"
UNIT = "GCUPS"
REF_BIN = ROOT / "oracle" / "_ref" / "SegAligner_ref"
N_SEGS, N_CONTIGS = 2000, 500


def pair_evals(n_labels: np.ndarray) -> np.ndarray:
    """sum_{j<n} min(j, 6) per map: the look-back window sizes of dp_align (SURVEY 8(d) work model)."""
    n = n_labels.astype(np.int64)
    full = np.maximum(n - 6, 0)
    head = np.minimum(n, 6)
    return full * 6 + head * (head - 1) // 2


def build_workload(rank: int, scale: float):
    from ampliconreconstructorom_b200 import capi, synth
    n_segs = max(int(round(N_SEGS * scale)), 4)
    n_contigs = max(int(round(N_CONTIGS * scale)), 4)
    segs, contigs = synth.make_workload(2000 + rank, n_segs, n_contigs)
    sid, soff, spos = capi.flatten_maps(segs)
    # both orientations (make_reverse_cmap) + non-collapse probabilities, with the executable's own host code
    fid, foff, fpos = capi.host_make_reverse(sid, soff, spos, 12)
    cid, coff, cpos = capi.flatten_maps(contigs)
    fprob = capi.host_non_collapse_probs(foff, fpos, 2)
    cprob = capi.host_non_collapse_probs(coff, cpos, 2)
    return dict(segs=segs, contigs=contigs, fid=fid, foff=foff, fpos=fpos, fprob=fprob, cid=cid, coff=coff, cpos=cpos,
                cprob=cprob, n_segs=n_segs, n_contigs=n_contigs)


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


def time_reference(workdir: Path, segs: dict, contigs: dict, n_threads: int):
    """One run of the unmodified reference binary; returns (wall seconds, scoring cells)."""
    from ampliconreconstructorom_b200 import synth
    workdir.mkdir(parents=True, exist_ok=True)
    synth.write_cmap(workdir / "segs.cmap", segs)
    synth.write_cmap(workdir / "contigs.cmap", contigs)
    nx = sum(len(m) - 1 for m in segs.values() if len(m) - 1 >= 12)  # make_reverse_cmap keeps labels >= min_labs
    nb = sum(len(m) - 1 for m in contigs.values())
    cells = 2.0 * nx * nb
    t0 = time.perf_counter()
    subprocess.run([str(REF_BIN), str(workdir / "segs.cmap"), str(workdir / "contigs.cmap"),
                    f"-prefix={workdir}/SA", f"-nthreads={n_threads}", "-min_labs=12", "-gen=2"],
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
    return time.perf_counter() - t0, cells


def reference_sample(w, target_s: float, n_threads: int):
    """A subsample of the rank-0 workload sized so that the reference needs about target_s seconds."""
    seg_ids = sorted(w["segs"])
    con_ids = sorted(w["contigs"])
    # probe: 4 segments x 32 contigs (>= 31 contigs keeps the thresholds defined)
    segs = {i: w["segs"][i] for i in seg_ids[:4]}
    contigs = {i: w["contigs"][i] for i in con_ids[:32]}
    with tempfile.TemporaryDirectory() as d:
        dt, cells = time_reference(Path(d), segs, contigs, n_threads)
    rate = cells / max(dt, 1e-3)
    n_con = min(len(con_ids), max(32, 4 * n_threads))
    contigs = {i: w["contigs"][i] for i in con_ids[:n_con]}
    nb = sum(len(m) - 1 for m in contigs.values())
    want_nx = rate * target_s / (2.0 * nb)
    segs, acc = {}, 0
    for i in seg_ids:
        segs[i] = w["segs"][i]
        acc += len(w["segs"][i]) - 1
        if acc >= want_nx:
            break
    return segs, contigs


def time_executable(w, frac: float):
    """The drop-in executable on a subsample of the rank-0 workload, whole run (parse, scoring, thresholds, alignment,
    tip alignment, output files) -- the same kind of measurement as the reference arm; returns a dict for the JSON."""
    from ampliconreconstructorom_b200 import capi, synth
    seg_ids, con_ids = sorted(w["segs"]), sorted(w["contigs"])
    n_s = max(4, int(len(seg_ids) * frac))
    n_c = max(40, int(len(con_ids) * frac))
    segs = {i: w["segs"][i] for i in seg_ids[:n_s]}
    contigs = {i: w["contigs"][i] for i in con_ids[:n_c]}
    nx = sum(len(m) - 1 for m in segs.values() if len(m) - 1 >= 12)
    nb = sum(len(m) - 1 for m in contigs.values())
    cells = 2.0 * nx * nb
    with tempfile.TemporaryDirectory() as d:
        d = Path(d)
        synth.write_cmap(d / "segs.cmap", segs)
        synth.write_cmap(d / "contigs.cmap", contigs)
        best = None
        for rep in range(2):  # first run pays the file-system and CUDA start-up warm-up
            t0 = time.perf_counter()
            p = subprocess.run([str(capi.BIN_PATH), str(d / "segs.cmap"), str(d / "contigs.cmap"), f"-prefix={d}/SA_r{rep}",
                                f"-nthreads={os.cpu_count() or 1}", "-min_labs=12", "-gen=2", "-ngpus=1"],
                               stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
            dt = time.perf_counter() - t0
            if p.returncode != 0:
                return {"value": None, "unit": UNIT, "error": p.stderr[-300:]}
            best = dt if best is None else min(best, dt)
        n_files = len([f for f in d.iterdir() if f.name.endswith("_aln.txt")])
    return {"value": cells / best / 1e9, "unit": UNIT, "wall_s": best, "scoring_cells": cells,
            "sample": f"bin/SegAligner, whole process (CUDA start-up, parse, phase-1 scoring, thresholds, alignment, tip "
                      f"alignment, {n_files // 2} alignment files) on {n_s} segments x {n_c} contigs of this workload, 1 GPU; "
                      "metric = phase-1 scoring cells / process wall time, the same definition as cpu_baseline"}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    base = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64",
            "data": "synthetic"}
    if not REF_BIN.exists():
        emit_json({"impl": "reference", "unavailable": "oracle/_ref/SegAligner_ref was not built"})
        return
    n_threads = os.cpu_count() or 1
    w = build_workload(0, args.scale)
    segs, contigs = reference_sample(w, args.ref_seconds, n_threads)
    times, cells = [], 0.0
    with tempfile.TemporaryDirectory() as d:
        for it in range(args.warmup_ref + args.steps):
            dt, cells = time_reference(Path(d) / f"r{it}", segs, contigs, n_threads)
            if it >= args.warmup_ref:
                times.append(dt)
    t = sum(times)
    gcups = cells * len(times) / t / 1e9
    n_prob = 2 * sum(1 for m in segs.values() if len(m) - 1 >= 12) * len(contigs)
    sample = (f"unmodified reference SegAligner, whole run (parse+score+thresholds+align) on {len(segs)} of "
              f"{w['n_segs']} segments x {len(contigs)} of {w['n_contigs']} contigs of the C2 workload, "
              f"{cells:.3e} scoring cells per step, -nthreads={n_threads}")
    base.update({"value": gcups, "ms_per_step": 1e3 * t / len(times), "alignments_per_s": n_prob * len(times) / t,
                 "config": {"workload": "C2 synthetic batch: 2,000 segments x 500 DLE1 contigs, both orientations "
                                        "(bounded sample, see cpu_baseline.sample)", "seed": 2000},
                 "cpu_baseline": {"value": gcups, "unit": UNIT, "cores": n_threads, "kind": "reference", "sample": sample},
                 "e2e": {"value": gcups, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0})
    emit_json(base)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=float(os.environ.get("SA_BENCH_SCALE", "1.0")),
                    help="linear scale of both map counts (1.0 = the named C2 config; <1 only for dry runs)")
    ap.add_argument("--ref-seconds", type=float, default=12.0, help="CPU seconds per reference step / baseline sample")
    ap.add_argument("--warmup-ref", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from ampliconreconstructorom_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = capi.Context(local_rank)
    w = build_workload(rank, args.scale)

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()

    keep = []
    bufs = {}
    for k in ("foff", "fpos", "fprob", "coff", "cpos", "cprob"):
        t, v = pinned(w[k])
        keep.append(t)
        bufs[k] = v
    n_c, n_s = len(w["cid"]), len(w["fid"])
    cc, ss = np.meshgrid(np.arange(n_c, dtype=np.int32), np.arange(n_s, dtype=np.int32), indexing="ij")
    problems = ctx.problems(cc.ravel(), ss.ravel())
    tp, problems = pinned(problems.view(np.uint8))
    problems = problems.view(capi.problem_dtype)
    keep.append(tp)
    n_prob = problems.size
    nb = (np.diff(w["coff"]) - 1)
    nx = (np.diff(w["foff"]) - 1)
    cells = float(nb.sum()) * float(nx.sum())
    evals = float(pair_evals(nb).sum()) * float(pair_evals(nx).sum())
    mode = ctx.mode(capi.SA_INIT_SEMIGLOBAL, capi.SA_START_SEMIGLOBAL, 0, 6, 0)

    # ---- resident inputs for the device-timed number
    ctx.set_maps(capi.SA_SIDE_CONTIG, bufs["coff"], bufs["cpos"], bufs["cprob"])
    ctx.set_maps(capi.SA_SIDE_SEG, bufs["foff"], bufs["fpos"], bufs["fprob"])
    plan = ctx.plan_create(problems)
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    peak = ctx.fp64_peak()

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
    checksum = float(np.sum(res["score"].astype(np.float64)))

    # ---- end to end through the host-buffer C-ABI calls (pinned host memory in, host results out)
    t_e2e = 0.0
    e2e_steps = max(1, min(args.steps, 2))
    for it in range(e2e_steps + 1):  # first pass untimed
        barrier()
        t1 = time.perf_counter()
        ctx.set_maps(capi.SA_SIDE_CONTIG, bufs["coff"], bufs["cpos"], bufs["cprob"])
        ctx.set_maps(capi.SA_SIDE_SEG, bufs["foff"], bufs["fpos"], bufs["fprob"])
        res2 = ctx.score_batch(mode, problems)
        torch.cuda.synchronize()
        if it > 0:
            t_e2e += time.perf_counter() - t1
    assert np.array_equal(res2["score"].view(np.uint32), res["score"].view(np.uint32)), "e2e and resident results differ"
    n_lab = int(nb.sum() + nx.sum())
    h2d = int(sum(bufs[k].nbytes for k in bufs) + 48 * n_lab + problems.nbytes + 4 * n_prob)  # + LabelOps + order
    d2h = int(res2.nbytes)

    # ---- reductions over ranks (max time, summed work)
    from ampliconreconstructorom_b200 import dist_util
    (dev_ms_max, wall_ms_max, e2e_ms_max, fill_sum_max), (cells_all, prob_all, evals_all, launches_all) = \
        dist_util.reduce_over_ranks([dev_ms, wall * 1e3, t_e2e * 1e3 / e2e_steps, sum(fill_ms)],
                                    [cells, float(n_prob), evals, float(launches)], device="cuda")

    if rank == 0:
        K = args.steps
        value = cells_all * K / (dev_ms_max * 1e-3) / 1e9
        e2e_val = cells_all / (e2e_ms_max * 1e-3) / 1e9
        achieved = 18.0 * evals * K / (sum(fill_ms) * 1e-3)  # this rank's kernel: fp64 thread-instructions / s
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 scores, f64 inside powf (bit-exact glibc replay)", "data": "synthetic",
            "config": {"workload": "C2 synthetic batch: 2,000 segments x 500 Saphyr DLE1 contigs, both orientations "
                                   "(phase-1 scoring: semi-global init, lookback-6 dp_align, backtrack-start scan)",
                       "scale": args.scale, "segments_both_orientations": int(n_s), "contigs": int(n_c),
                       "problems_per_gpu": int(n_prob), "cells_per_gpu": cells, "seed": "2000+rank",
                       "l2": "flushed between steps (256 MiB write)", "sharding": "independent batch per rank, no collective",
                       "large_problems_cooperative": info["n_large"]},
            "alignments_per_s": prob_all * K / (dev_ms_max * 1e-3),
            "wall_ms_per_step": wall_ms_max / K,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "alignments_per_s": prob_all / (e2e_ms_max * 1e-3), "steps": e2e_steps},
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "fp64_issue", "achieved": achieved / 1e12, "peak": peak / 1e12,
                         "unit": "T fp64-instr/s (thread-level DFMA/DADD/DMUL)", "frac": achieved / peak,
                         "peak_source": "live DFMA microbenchmark on this GPU (sa_measure_fp64_peak); "
                                        "MEASURED_PEAKS.json has no fp64 entry",
                         "algorithmic": "18 fp64 instr per predecessor evaluation x pair evaluations per launch "
                                        "(the reference's un-pruned work, SURVEY 8(d)); frac > 1 is possible because the "
                                        "kernel prunes exactly: fp32 bounds for all candidates, one bit-exact replay per cell",
                         "executed": {"powf_replays_per_launch": n_replays, "cells_rescanned_per_launch": n_rescans,
                                      "replays_per_cell": n_replays / cells,
                                      "fp64_instr_per_s_T": 17.0 * n_replays / (fill_ms[-1] * 1e-3) / 1e12,
                                      "frac_of_fp64_peak": 17.0 * n_replays / (fill_ms[-1] * 1e-3) / peak,
                                      # XU pipe: 2 MUFU per predecessor evaluation (bound pass) + 2 F2F per replay, against the
                                      # nominal 16 lanes/clk/SM x 148 SMs x the SM clock sampled during the run
                                      "xu_instr_per_s_T": (2.0 * evals + 2.0 * n_replays) / (fill_ms[-1] * 1e-3) / 1e12,
                                      "xu_frac_of_nominal": (2.0 * evals + 2.0 * n_replays) / (fill_ms[-1] * 1e-3)
                                      / (16.0 * 148 * 1e6 * max(float(summarise_clocks(samples).get("sm_mhz") or 1965.0), 1.0)),
                                      "note": "executed work is bound by the XU pipe (2 MUFU per candidate, 16 lanes/clk/SM) "
                                              "and by issue slots, not by fp64 issue; see DESIGN.md 4.1"},
                         "kernel_ms_per_launch": sum(fill_ms) / K, "traffic": None,
                         "hbm_note": "HBM traffic is O(results): <1e-3 of measured copy bandwidth"},
            "clocks": summarise_clocks(samples),
            "checksum": checksum,
        }
        if world == 1:
            if args.no_cpu_baseline:
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "skipped (--no-cpu-baseline)"}
            elif REF_BIN.exists():
                n_threads = os.cpu_count() or 1
                segs, contigs = reference_sample(w, args.ref_seconds, n_threads)
                with tempfile.TemporaryDirectory() as d:
                    dt, c = time_reference(Path(d), segs, contigs, n_threads)
                out["cpu_baseline"] = {"value": c / dt / 1e9, "unit": UNIT, "cores": n_threads, "kind": "reference",
                                       "sample": f"unmodified reference SegAligner (oracle/_ref), whole run on {len(segs)} segments x "
                                                 f"{len(contigs)} contigs of this workload, {c:.3e} scoring cells in {dt:.1f} s wall, "
                                                 f"-nthreads={n_threads}"}
            else:
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                       "sample": "oracle/_ref/SegAligner_ref not present"}
            if not args.no_cpu_baseline:
                # the same whole-run definition as cpu_baseline, for the drop-in executable (all phases, not only the kernel)
                out["whole_run"] = time_executable(w, 0.3)
        emit_json(out)
    plan.destroy()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
