"""ctypes binding of libsegaligner_b200.so (C ABI in include/segaligner_b200.h).

There is no CPU fallback: importing works without a GPU (so the symbol table can be checked), but creating a
Context raises if the shared library is missing or no sm_100 device is usable.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("SA_LIB_PATH", PKG_DIR / "lib" / "libsegaligner_b200.so"))  # override: dev experiments only
BIN_PATH = PKG_DIR / "bin" / "SegAligner"

SA_INIT_SEMIGLOBAL, SA_INIT_LOCAL, SA_INIT_FITTING = 0, 1, 2
SA_START_SEMIGLOBAL, SA_START_LOCAL, SA_START_END = 0, 1, 2
SA_SIDE_CONTIG, SA_SIDE_SEG = 0, 1

problem_dtype = np.dtype([("contig", "<i4"), ("seg", "<i4"), ("used_slot", "<i4"), ("reserved", "<i4")])
result_dtype = np.dtype([("score", "<f4"), ("j", "<i4"), ("q", "<i4"), ("path_len", "<i4")])


class Mode(C.Structure):
    _fields_ = [("init_mode", C.c_int32), ("start_mode", C.c_int32), ("fitting", C.c_int32),
                ("lookback", C.c_int32), ("swap_b_x", C.c_int32)]


class PackedPaths(C.Structure):
    _fields_ = [("n", C.c_int64), ("total", C.c_int64), ("results", C.c_void_p), ("start", C.c_void_p),
                ("path_j", C.c_void_p), ("path_q", C.c_void_p), ("path_s", C.c_void_p)]


class RoundsOut(C.Structure):
    _fields_ = [("n", C.c_int64), ("max_alns", C.c_int32), ("reserved", C.c_int32), ("total", C.c_int64),
                ("n_rounds", C.c_void_p), ("rounds", C.c_void_p), ("path_j", C.c_void_p), ("path_q", C.c_void_p),
                ("path_s", C.c_void_p), ("round_off", C.c_void_p)]


class RoundsFilter(C.Structure):
    """sa_rounds_filter: which tracebacks sa_align_rounds_filtered returns (main.cpp:303-325)."""
    _fields_ = [("tip", C.c_int32), ("min_len", C.c_int32), ("mean_min", C.c_float), ("mean_min_short", C.c_float),
                ("stop_below", C.c_float), ("tip_windows", C.c_int32)]


round_dtype = np.dtype([("score", "<f4"), ("j", "<i4"), ("q", "<i4"), ("path_len", "<i4"), ("path_off", "<i8"), ("j_first", "<i4"), ("reserved", "<i4")])


class Timing(C.Structure):
    _fields_ = [("fill_ms", C.c_float), ("post_ms", C.c_float), ("n_launches", C.c_int32), ("reserved", C.c_int32)]


# every symbol include/segaligner_b200.h declares (tests check the library exports exactly these)
EXPORTS = [
    "sa_device_count", "sa_ctx_create", "sa_ctx_destroy", "sa_last_error", "sa_set_maps", "sa_score_batch",
    "sa_plan_create", "sa_plan_destroy", "sa_plan_score", "sa_plan_results", "sa_plan_info", "sa_sync", "sa_align_batch", "sa_align_batch_packed", "sa_align_rounds", "sa_align_rounds_filtered", "sa_prepare_rounds", "sa_detect_batch", "sa_fill_one", "sa_powf12",
    "sa_powf12_range", "sa_check_prune_bound", "sa_get_timing", "sa_get_prune_stats", "sa_measure_fp64_peak", "sa_measure_mufu_peak", "sa_get_traffic", "sa_event_record", "sa_event_elapsed_ms",
    "sa_host_non_collapse_probs", "sa_host_make_reverse", "sa_host_parse_cmap", "sa_host_score_thresholds",
    "sa_host_format_float", "sa_digest",
]

_lib = None


def load_library() -> C.CDLL:
    """dlopen the in-tree shared library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the SegAligner DP path)")
    lib = C.CDLL(str(LIB_PATH))
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.sa_device_count.restype = C.c_int
    lib.sa_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.sa_ctx_destroy.argtypes = [vp]
    lib.sa_ctx_destroy.restype = None
    lib.sa_last_error.argtypes = [vp]
    lib.sa_last_error.restype = C.c_char_p
    lib.sa_set_maps.argtypes = [vp, C.c_int, i32, vp, vp, vp]
    lib.sa_score_batch.argtypes = [vp, C.POINTER(Mode), vp, i64, vp]
    lib.sa_plan_create.argtypes = [vp, vp, i64, C.POINTER(vp)]
    lib.sa_plan_destroy.argtypes = [vp, vp]
    lib.sa_plan_destroy.restype = None
    lib.sa_plan_score.argtypes = [vp, vp, C.POINTER(Mode)]
    lib.sa_plan_results.argtypes = [vp, vp, vp]
    lib.sa_plan_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(C.c_double)]
    lib.sa_sync.argtypes = [vp]
    lib.sa_align_batch.argtypes = [vp, C.POINTER(Mode), vp, i64, vp, i64, i64, vp, vp, vp, vp, vp]
    lib.sa_align_batch_packed.argtypes = [vp, C.POINTER(Mode), vp, i64, vp, i64, i64, C.POINTER(PackedPaths)]
    lib.sa_align_rounds.argtypes = [vp, C.POINTER(Mode), vp, i64, vp, i32, vp, i64, i64, C.POINTER(RoundsOut)]
    lib.sa_align_rounds_filtered.argtypes = [vp, C.POINTER(Mode), vp, i64, vp, i32, vp, i64, i64, C.POINTER(RoundsFilter), C.POINTER(RoundsOut)]
    lib.sa_prepare_rounds.argtypes = [vp, i64, i32]
    lib.sa_detect_batch.argtypes = [vp, C.POINTER(Mode), vp, i64, vp, vp, vp, vp]
    lib.sa_fill_one.argtypes = [vp, C.POINTER(Mode), vp, vp, i64, vp, vp, vp]
    lib.sa_powf12.argtypes = [vp, vp, i64, vp, C.c_int]
    lib.sa_powf12_range.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp]
    lib.sa_check_prune_bound.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_float)]
    lib.sa_get_timing.argtypes = [vp, C.POINTER(Timing)]
    lib.sa_get_prune_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.sa_measure_fp64_peak.argtypes = [vp, C.POINTER(C.c_double)]
    lib.sa_measure_mufu_peak.argtypes = [vp, C.POINTER(C.c_double)]
    lib.sa_get_traffic.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.sa_digest.argtypes = [vp, vp, i64, C.c_char_p, i32, i32, i64, vp, C.POINTER(i64)]
    lib.sa_event_record.argtypes = [vp, i32]
    lib.sa_event_elapsed_ms.argtypes = [vp, C.POINTER(C.c_float)]
    lib.sa_host_non_collapse_probs.argtypes = [i32, vp, vp, i32, vp]
    lib.sa_host_make_reverse.argtypes = [i32, vp, vp, vp, i32, i32, i64, vp, vp, vp, C.POINTER(i32), C.POINTER(i64)]
    lib.sa_host_parse_cmap.argtypes = [C.c_char_p, i32, i64, vp, vp, vp, C.POINTER(i32), C.POINTER(i64)]
    lib.sa_host_score_thresholds.argtypes = [i64, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, C.c_double, i32, vp]
    lib.sa_host_format_float.argtypes = [C.c_float, C.c_char_p, i32]
    _lib = lib
    return lib


# ---- host-side (CPU) preparation helpers: the executable's own C++ behind the C ABI -------------------------------

def flatten_maps(maps: dict):
    """{id: positions incl. length} -> (ids int32, off int64, pos float32) in ascending-id order."""
    ids = np.array(sorted(maps), dtype=np.int32)
    off = np.zeros(ids.size + 1, dtype=np.int64)
    for k, i in enumerate(ids):
        off[k + 1] = off[k] + len(maps[int(i)])
    pos = (np.concatenate([np.asarray(maps[int(i)], dtype=np.float32) for i in ids])
           if ids.size else np.zeros(0, dtype=np.float32))
    return ids, off, pos


def unflatten_maps(ids, off, pos) -> dict:
    return {int(ids[k]): pos[off[k]:off[k + 1]].copy() for k in range(len(ids))}


def host_non_collapse_probs(off, pos, gen: int) -> np.ndarray:
    lib = load_library()
    off = np.ascontiguousarray(off, dtype=np.int64)
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    n = off.size - 1
    out = np.zeros(max(int(off[-1]) - n, 1), dtype=np.float32)
    rc = lib.sa_host_non_collapse_probs(n, _ptr(off), _ptr(pos), gen, _ptr(out))
    if rc:
        raise SegAlignerError(f"sa_host_non_collapse_probs failed ({rc})")
    return out[:int(off[-1]) - n]


def host_make_reverse(ids, off, pos, min_map_len: int):
    lib = load_library()
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    off = np.ascontiguousarray(off, dtype=np.int64)
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    n_out, n_pos = C.c_int32(), C.c_int64()
    lib.sa_host_make_reverse(ids.size, _ptr(ids), _ptr(off), _ptr(pos), min_map_len, 0, 0, None, None, None,
                             C.byref(n_out), C.byref(n_pos))
    oi = np.zeros(max(n_out.value, 1), dtype=np.int32)
    oo = np.zeros(n_out.value + 1, dtype=np.int64)
    op = np.zeros(max(n_pos.value, 1), dtype=np.float32)
    rc = lib.sa_host_make_reverse(ids.size, _ptr(ids), _ptr(off), _ptr(pos), min_map_len, n_out.value, n_pos.value,
                                  _ptr(oi), _ptr(oo), _ptr(op), C.byref(n_out), C.byref(n_pos))
    if rc:
        raise SegAlignerError(f"sa_host_make_reverse failed ({rc})")
    return oi[:n_out.value], oo, op[:n_pos.value]


def host_parse_cmap(path):
    lib = load_library()
    n_out, n_pos = C.c_int32(), C.c_int64()
    rc = lib.sa_host_parse_cmap(str(path).encode(), 0, 0, None, None, None, C.byref(n_out), C.byref(n_pos))
    if rc:
        raise SegAlignerError(f"sa_host_parse_cmap({path}) failed ({rc})")
    oi = np.zeros(max(n_out.value, 1), dtype=np.int32)
    oo = np.zeros(n_out.value + 1, dtype=np.int64)
    op = np.zeros(max(n_pos.value, 1), dtype=np.float32)
    lib.sa_host_parse_cmap(str(path).encode(), n_out.value, n_pos.value, _ptr(oi), _ptr(oo), _ptr(op),
                           C.byref(n_out), C.byref(n_pos))
    return oi[:n_out.value], oo, op[:n_pos.value]


def host_score_thresholds(row_seg, row_contig, row_score, segs: dict, contigs: dict, pvalue: float, n_detect=500):
    lib = load_library()
    si, so, sp = flatten_maps(segs)
    ci, co, cp = flatten_maps(contigs)
    row_seg = np.ascontiguousarray(row_seg, dtype=np.int32)
    row_contig = np.ascontiguousarray(row_contig, dtype=np.int32)
    row_score = np.ascontiguousarray(row_score, dtype=np.float32)
    out = np.zeros(max(si.size, 1), dtype=np.float32)
    rc = lib.sa_host_score_thresholds(row_score.size, _ptr(row_seg), _ptr(row_contig), _ptr(row_score), si.size,
                                      _ptr(si), _ptr(so), _ptr(sp), ci.size, _ptr(ci), _ptr(co), _ptr(cp),
                                      C.c_double(pvalue), n_detect, _ptr(out))
    if rc:
        raise SegAlignerError(f"sa_host_score_thresholds failed ({rc})")
    return {int(si[k]): out[k] for k in range(si.size)}


def host_format_float(v: float) -> str:
    lib = load_library()
    buf = C.create_string_buffer(64)
    lib.sa_host_format_float(C.c_float(v), buf, 64)
    return buf.value.decode()


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class SegAlignerError(RuntimeError):
    pass


class Plan:
    """Device-resident problem list (sa_plan)."""

    def __init__(self, ctx, h, n):
        self.ctx, self.h, self.n = ctx, h, n

    def score(self, mode: Mode):
        self.ctx._check(self.ctx.lib.sa_plan_score(self.ctx.h, self.h, C.byref(mode)), "sa_plan_score")

    def results(self) -> np.ndarray:
        res = np.zeros(self.n, dtype=result_dtype)
        self.ctx._check(self.ctx.lib.sa_plan_results(self.ctx.h, self.h, _ptr(res)), "sa_plan_results")
        return res

    def info(self):
        nb, nl, tc = C.c_int64(), C.c_int64(), C.c_double()
        self.ctx.lib.sa_plan_info(self.h, C.byref(nb), C.byref(nl), C.byref(tc))
        return {"n_bulk": nb.value, "n_large": nl.value, "total_cells": tc.value}

    def destroy(self):
        if self.h:
            self.ctx.lib.sa_plan_destroy(self.ctx.h, self.h)
            self.h = None


class Context:
    """One GPU context (sa_ctx). Not thread-safe; use one per host thread / device."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.sa_ctx_create(int(device), C.byref(h))
        if rc != 0:
            raise SegAlignerError(f"sa_ctx_create(device={device}) failed with {rc}: no usable sm_100 GPU "
                                  "(the DP path has no CPU fallback)")
        self.h = h
        self.n_labels = [None, None]

    def close(self):
        if getattr(self, "h", None):
            self.lib.sa_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise SegAlignerError(f"{what} failed ({rc}): {self.lib.sa_last_error(self.h).decode()}")

    def set_maps(self, side: int, off: np.ndarray, pos: np.ndarray, prob: np.ndarray):
        off = np.ascontiguousarray(off, dtype=np.int64)
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        prob = np.ascontiguousarray(prob, dtype=np.float32)
        n_maps = len(off) - 1
        assert pos.size == off[-1] and prob.size == off[-1] - n_maps
        self._check(self.lib.sa_set_maps(self.h, side, n_maps, _ptr(off), _ptr(pos), _ptr(prob)), "sa_set_maps")
        self.n_labels[side] = (np.diff(off) - 1).astype(np.int64)

    @staticmethod
    def mode(init_mode=SA_INIT_SEMIGLOBAL, start_mode=SA_START_SEMIGLOBAL, fitting=0, lookback=6, swap_b_x=0):
        return Mode(init_mode, start_mode, fitting, lookback, swap_b_x)

    @staticmethod
    def problems(contig, seg, used_slot=None):
        contig = np.asarray(contig, dtype=np.int32)
        pr = np.zeros(contig.size, dtype=problem_dtype)
        pr["contig"] = contig
        pr["seg"] = np.asarray(seg, dtype=np.int32)
        pr["used_slot"] = -1 if used_slot is None else np.asarray(used_slot, dtype=np.int32)
        return pr

    def score_batch(self, mode: Mode, problems: np.ndarray) -> np.ndarray:
        problems = np.ascontiguousarray(problems, dtype=problem_dtype)
        res = np.zeros(problems.size, dtype=result_dtype)
        self._check(self.lib.sa_score_batch(self.h, C.byref(mode), _ptr(problems), problems.size, _ptr(res)),
                    "sa_score_batch")
        return res

    # staged form of score_batch: problems stay resident in HBM between plan_score calls
    def plan_create(self, problems: np.ndarray):
        problems = np.ascontiguousarray(problems, dtype=problem_dtype)
        h = C.c_void_p()
        self._check(self.lib.sa_plan_create(self.h, _ptr(problems), problems.size, C.byref(h)), "sa_plan_create")
        return Plan(self, h, problems.size)

    def sync(self):
        self._check(self.lib.sa_sync(self.h), "sa_sync")

    def align_batch(self, mode: Mode, problems: np.ndarray, used_bits: np.ndarray | None = None,
                    used_words_per_slot: int = 0):
        """Returns (results, path_off, path_j, path_q, path_s); paths are end->start like get_aln_list."""
        problems = np.ascontiguousarray(problems, dtype=problem_dtype)
        n = problems.size
        nb = self.n_labels[0][problems["contig"]]
        nx = self.n_labels[1][problems["seg"]]
        path_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.minimum(nb, nx), out=path_off[1:])
        tot = max(int(path_off[-1]), 1)
        pj = np.full(tot, -1, dtype=np.int32)
        pq = np.full(tot, -1, dtype=np.int32)
        ps = np.zeros(tot, dtype=np.float32)
        res = np.zeros(n, dtype=result_dtype)
        n_slots = 0
        if used_bits is not None:
            used_bits = np.ascontiguousarray(used_bits, dtype=np.uint32)
            n_slots = used_bits.size // max(used_words_per_slot, 1)
        self._check(self.lib.sa_align_batch(self.h, C.byref(mode), _ptr(problems), n, _ptr(used_bits),
                                            used_words_per_slot, n_slots, _ptr(path_off), _ptr(res), _ptr(pj),
                                            _ptr(pq), _ptr(ps)), "sa_align_batch")
        return res, path_off, pj, pq, ps

    def align_batch_packed(self, mode: Mode, problems: np.ndarray, used_bits: np.ndarray | None = None,
                           used_words_per_slot: int = 0):
        """sa_align_batch_packed: returns copies (results, start, path_j, path_q, path_s) of the context-owned buffers;
        path k is path_*[start[k] : start[k] + results['path_len'][k]], end->start."""
        problems = np.ascontiguousarray(problems, dtype=problem_dtype)
        n = problems.size
        n_slots = 0
        if used_bits is not None:
            used_bits = np.ascontiguousarray(used_bits, dtype=np.uint32)
            n_slots = used_bits.size // max(used_words_per_slot, 1)
        pk = PackedPaths()
        self._check(self.lib.sa_align_batch_packed(self.h, C.byref(mode), _ptr(problems), n, _ptr(used_bits),
                                                   used_words_per_slot, n_slots, C.byref(pk)), "sa_align_batch_packed")

        def view(addr, count, dtype):
            if count == 0 or not addr:
                return np.zeros(0, dtype=dtype)
            buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(addr)
            return np.frombuffer(buf, dtype=dtype).copy()

        return (view(pk.results, n, result_dtype), view(pk.start, n + 1, np.int64), view(pk.path_j, pk.total, np.int32),
                view(pk.path_q, pk.total, np.int32), view(pk.path_s, pk.total, np.float32))

    def align_rounds(self, mode: Mode, problems: np.ndarray, thr: np.ndarray, max_alns: int,
                     used_bits: np.ndarray | None = None, used_words_per_slot: int = 0, flt: "RoundsFilter | None" = None):
        """sa_align_rounds / sa_align_rounds_filtered (flt given): the whole per-pair run_SA_aln loop on the device.  Returns
        copies (n_rounds[n], rounds[n, max_alns] (round_dtype), path_j, path_q, path_s)."""
        problems = np.ascontiguousarray(problems, dtype=problem_dtype)
        thr = np.ascontiguousarray(thr, dtype=np.float32)
        n = problems.size
        n_slots = 0
        if used_bits is not None:
            used_bits = np.ascontiguousarray(used_bits, dtype=np.uint32)
            n_slots = used_bits.size // max(used_words_per_slot, 1)
        ro = RoundsOut()
        if flt is None:
            self._check(self.lib.sa_align_rounds(self.h, C.byref(mode), _ptr(problems), n, _ptr(thr), max_alns, _ptr(used_bits),
                                                 used_words_per_slot, n_slots, C.byref(ro)), "sa_align_rounds")
        else:
            self._check(self.lib.sa_align_rounds_filtered(self.h, C.byref(mode), _ptr(problems), n, _ptr(thr), max_alns, _ptr(used_bits),
                                                          used_words_per_slot, n_slots, C.byref(flt), C.byref(ro)), "sa_align_rounds_filtered")

        def view(addr, count, dtype):
            if count == 0 or not addr:
                return np.zeros(0, dtype=dtype)
            buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(addr)
            return np.frombuffer(buf, dtype=dtype).copy()

        # the library returns the records packed (pair k: rounds[round_off[k] .. + n_rounds[k])); callers here index [pair, round]
        nr = view(ro.n_rounds, n, np.int32)
        off = view(ro.round_off, n + 1, np.int64)
        packed = view(ro.rounds, int(off[-1]) if n else 0, round_dtype)
        assert np.array_equal(np.diff(off), np.maximum(nr, 0)), "round_off is not the prefix sum of the round counts"
        dense = np.zeros((n, max_alns), dtype=round_dtype)
        for r in range(max_alns):
            has = nr > r
            dense[has, r] = packed[off[:-1][has] + r]
        return (nr, dense, view(ro.path_j, ro.total, np.int32), view(ro.path_q, ro.total, np.int32), view(ro.path_s, ro.total, np.float32))

    def detect_batch(self, mode: Mode, problems: np.ndarray, n_want: np.ndarray):
        problems = np.ascontiguousarray(problems, dtype=problem_dtype)
        n = problems.size
        n_want = np.ascontiguousarray(n_want, dtype=np.int32)
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(n_want, out=off[1:])
        scores = np.zeros(max(int(off[-1]), 1), dtype=np.float32)
        got = np.zeros(n, dtype=np.int32)
        self._check(self.lib.sa_detect_batch(self.h, C.byref(mode), _ptr(problems), n, _ptr(n_want), _ptr(off),
                                             _ptr(scores), _ptr(got)), "sa_detect_batch")
        return scores, off, got

    def fill_one(self, mode: Mode, contig: int, seg: int, used_rows=None):
        nb = int(self.n_labels[0][contig])
        nx = int(self.n_labels[1][seg])
        pr = self.problems([contig], [seg])
        S = np.zeros((nb, nx), dtype=np.float32)
        prev = np.zeros((nb, nx, 2), dtype=np.int32)
        res = np.zeros(1, dtype=result_dtype)
        used_bits, words = None, 0
        if used_rows is not None:
            words = (nb + 31) // 32
            used_bits = np.zeros(words, dtype=np.uint32)
            for j in used_rows:
                used_bits[j >> 5] |= np.uint32(1 << (j & 31))
        self._check(self.lib.sa_fill_one(self.h, C.byref(mode), _ptr(pr), _ptr(used_bits), words, _ptr(S),
                                         _ptr(prev), _ptr(res)), "sa_fill_one")
        return S, prev, res[0]

    def powf12(self, x: np.ndarray, fast: bool = True) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.zeros_like(x)
        self._check(self.lib.sa_powf12(self.h, _ptr(x), x.size, _ptr(out), 1 if fast else 0), "sa_powf12")
        return out

    def prune_stats(self):
        """(bit-exact powf replays, cells rescanned exactly) executed by the fill kernels of the last batch call."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._check(self.lib.sa_get_prune_stats(self.h, C.byref(a), C.byref(b)), "sa_get_prune_stats")
        return int(a.value), int(b.value)

    def check_prune_bound(self, lo_bits: int, hi_bits: int):
        """(violations, min exact/bound) of the pruning lower bound over the float bit patterns [lo_bits, hi_bits)."""
        n, r = C.c_uint64(0), C.c_float(0)
        self._check(self.lib.sa_check_prune_bound(self.h, lo_bits, hi_bits, C.byref(n), C.byref(r)), "sa_check_prune_bound")
        return int(n.value), float(r.value)

    def powf12_range(self, lo_bits: int, hi_bits: int):
        n = hi_bits - lo_bits
        a = np.empty(n, dtype=np.float32)
        b = np.empty(n, dtype=np.float32)
        self._check(self.lib.sa_powf12_range(self.h, lo_bits, hi_bits, _ptr(a), _ptr(b)), "sa_powf12_range")
        return a, b

    def timing(self) -> Timing:
        t = Timing()
        self._check(self.lib.sa_get_timing(self.h, C.byref(t)), "sa_get_timing")
        return t

    def digest(self, seq, motif: str, offset: int) -> np.ndarray:
        """In-silico digestion (generate_cmap.py:119-123): ascending unique label positions (int64) of `motif` and its
        reverse complement in `seq` (bytes / uint8 array); offset = the enzyme's cut offset + 1."""
        buf = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray, memoryview)) else \
            np.ascontiguousarray(seq, dtype=np.uint8)
        n_out = C.c_int64()
        m = motif.encode()
        self._check(self.lib.sa_digest(self.h, _ptr(buf) if buf.size else None, buf.size, m, len(m), offset, 0, None,
                                       C.byref(n_out)), "sa_digest") if buf.size else None
        if buf.size == 0 or n_out.value == 0:
            return np.zeros(0, dtype=np.int64)
        out = np.zeros(n_out.value, dtype=np.int64)
        self._check(self.lib.sa_digest(self.h, _ptr(buf), buf.size, m, len(m), offset, out.size, _ptr(out),
                                       C.byref(n_out)), "sa_digest")
        return out[:n_out.value]

    def event_record(self, slot: int):
        self._check(self.lib.sa_event_record(self.h, slot), "sa_event_record")

    def event_elapsed_ms(self) -> float:
        ms = C.c_float()
        self._check(self.lib.sa_event_elapsed_ms(self.h, C.byref(ms)), "sa_event_elapsed_ms")
        return ms.value

    def fp64_peak(self) -> float:
        v = C.c_double()
        self._check(self.lib.sa_measure_fp64_peak(self.h, C.byref(v)), "sa_measure_fp64_peak")
        return v.value

    def traffic(self):
        """(h2d bytes, d2h bytes) copied by this context so far."""
        a, b = C.c_uint64(), C.c_uint64()
        self._check(self.lib.sa_get_traffic(self.h, C.byref(a), C.byref(b)), "sa_get_traffic")
        return a.value, b.value

    def mufu_peak(self) -> float:
        v = C.c_double()
        self._check(self.lib.sa_measure_mufu_peak(self.h, C.byref(v)), "sa_measure_mufu_peak")
        return v.value
