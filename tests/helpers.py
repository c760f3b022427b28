"""Test-side loaders for the checkers under oracle/ (TEST INFRASTRUCTURE; never imported by the product)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
REF_BIN = ORACLE_DIR / "_ref" / "SegAligner_ref"
GOLDEN = Path(__file__).resolve().parent / "golden"


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def ensure_oracle_built():
    so = ORACLE_DIR / "libsa_oracle.so"
    if not so.exists():
        subprocess.check_call(["make", "-C", str(ORACLE_DIR), "libsa_oracle.so"], stdout=subprocess.DEVNULL)
    return so


class Oracle:
    """oracle/sa_oracle.c -- plain-C restatement of the reference hot path."""

    def __init__(self):
        self.lib = C.CDLL(str(ensure_oracle_built()))
        L = self.lib
        L.sao_powf12.restype = C.c_float
        L.sao_powf12.argtypes = [C.c_float]
        L.sao_score_f.restype = C.c_float
        L.sao_score_f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.sao_score_problem.restype = C.c_float
        L.sao_score_threshold.restype = C.c_float
        L.sao_score_threshold.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
        L.sao_multi_backtrack.restype = C.c_int
        L.sao_traceback.restype = C.c_int

    def probs(self, maps: dict, gen: int) -> dict:
        """non_collapse_probs over a whole map set in ascending-id order (carry crosses maps)."""
        carry = C.c_float(1.0)
        out = {}
        for mid in sorted(maps):
            pos = np.ascontiguousarray(maps[mid], dtype=np.float32)
            o = np.zeros(max(pos.size - 1, 0), dtype=np.float32)
            self.lib.sao_non_collapse_probs(_p(pos), C.c_int(pos.size), C.c_int(gen), C.byref(carry), _p(o))
            out[mid] = o
        return out

    def reverse(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        o = np.zeros_like(pos)
        self.lib.sao_reverse_map(_p(pos), C.c_int(pos.size), _p(o))
        return o

    def score_f(self, b, x, i, j, p, q, probs):
        b = np.ascontiguousarray(b, dtype=np.float32)
        x = np.ascontiguousarray(x, dtype=np.float32)
        probs = np.ascontiguousarray(probs, dtype=np.float32)
        return self.lib.sao_score_f(_p(b), _p(x), i, j, p, q, _p(probs))

    def fill(self, b, x, bprob, xprob, init_mode=0, fitting=0, lookback=6, used=None, swap=0):
        b = np.ascontiguousarray(b, dtype=np.float32)
        x = np.ascontiguousarray(x, dtype=np.float32)
        bprob = np.ascontiguousarray(bprob, dtype=np.float32)
        xprob = np.ascontiguousarray(xprob, dtype=np.float32)
        nb, nx = b.size - 1, x.size - 1
        S = np.zeros((nb, nx), dtype=np.float32)
        prev = np.zeros((nb, nx, 2), dtype=np.int32)
        self.lib.sao_init(_p(S), C.c_int(nb), C.c_int(nx), C.c_int(init_mode))
        u = None
        if used is not None:
            u = np.zeros(nb, dtype=np.uint8)
            u[list(used)] = 1
        self.lib.sao_dp_align(_p(S), _p(prev), _p(b), C.c_int(nb), _p(x), C.c_int(nx), C.c_int(fitting),
                              C.c_int(lookback), _p(u), _p(xprob), _p(bprob), C.c_int(swap))
        return S, prev

    def start(self, S, mode):
        nb, nx = S.shape
        jq = np.zeros(2, dtype=np.int32)
        if mode == 0:
            self.lib.sao_start_sg(_p(S), C.c_int(nb), C.c_int(nx), _p(jq))
        elif mode == 1:
            self.lib.sao_start_local(_p(S), C.c_int(nb), C.c_int(nx), _p(jq))
        else:
            jq[:] = (nb - 1, nx - 1)
        return int(jq[0]), int(jq[1])

    def traceback(self, S, prev, j, q):
        nb, nx = S.shape
        cap = min(nb, nx) + 1
        pjq = np.zeros((cap, 2), dtype=np.int32)
        ps = np.zeros(cap, dtype=np.float32)
        n = self.lib.sao_traceback(_p(S), _p(prev), C.c_int(nx), C.c_int(j), C.c_int(q), _p(pjq), _p(ps))
        return pjq[:n, 0].copy(), pjq[:n, 1].copy(), ps[:n].copy()

    def multi_backtrack(self, S, prev, n):
        S = S.copy()
        nb, nx = S.shape
        out = np.zeros(max(n, 1), dtype=np.float32)
        got = self.lib.sao_multi_backtrack(_p(S), _p(prev), C.c_int(nb), C.c_int(nx), C.c_int(n), _p(out))
        return out[:got].copy()

    def threshold(self, scores, n_contigs, n_detect, m_total, seg_labels, pvalue):
        scores = np.ascontiguousarray(scores, dtype=np.float32)
        return self.lib.sao_score_threshold(_p(scores), scores.size, n_contigs, n_detect, m_total, seg_labels,
                                            C.c_double(pvalue))


class RefShim:
    """oracle/_ref/libsa_ref.so -- the UNMODIFIED reference functions behind extern "C" wrappers."""

    @staticmethod
    def try_load():
        so = ORACLE_DIR / "_ref" / "libsa_ref.so"
        if not so.exists():
            return None
        return RefShim(so)

    def __init__(self, so):
        self.lib = C.CDLL(str(so))
        self.lib.ref_score_f.restype = C.c_float
        self.lib.ref_score_f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_void_p, C.c_int]
        self.lib.ref_dp_align.restype = C.c_int
        self.lib.ref_multi_backtrack.restype = C.c_int
        self.lib.ref_make_reverse_cmap.restype = C.c_int
        self.lib.ref_parse_cmap.restype = C.c_int64

    def probs(self, maps: dict, gen: int) -> dict:
        ids = np.array(sorted(maps), dtype=np.int32)
        off = np.zeros(ids.size + 1, dtype=np.int64)
        for k, i in enumerate(ids):
            off[k + 1] = off[k] + len(maps[i])
        pos = np.concatenate([np.asarray(maps[i], dtype=np.float32) for i in ids])
        out = np.zeros(int(off[-1]) - ids.size, dtype=np.float32)
        self.lib.ref_non_collapse_probs(C.c_int(ids.size), _p(ids), _p(off), _p(pos), C.c_int(gen), _p(out))
        res, o = {}, 0
        for k, i in enumerate(ids):
            n = len(maps[i]) - 1
            res[int(i)] = out[o:o + n].copy()
            o += n
        return res

    def make_reverse(self, maps: dict, min_map_len: int) -> dict:
        ids = np.array(sorted(maps), dtype=np.int32)
        off = np.zeros(ids.size + 1, dtype=np.int64)
        for k, i in enumerate(ids):
            off[k + 1] = off[k] + len(maps[i])
        pos = np.concatenate([np.asarray(maps[i], dtype=np.float32) for i in ids])
        oi = np.zeros(2 * ids.size + 1, dtype=np.int32)
        oo = np.zeros(2 * ids.size + 2, dtype=np.int64)
        op = np.zeros(2 * pos.size + 1, dtype=np.float32)
        n = self.lib.ref_make_reverse_cmap(C.c_int(ids.size), _p(ids), _p(off), _p(pos), C.c_int(min_map_len),
                                           _p(oi), _p(oo), _p(op))
        return {int(oi[k]): op[oo[k]:oo[k + 1]].copy() for k in range(n)}

    def parse_cmap(self, path) -> dict:
        n_maps = C.c_int(0)
        tot = self.lib.ref_parse_cmap(str(path).encode(), C.c_int(0), C.c_int64(0), None, None, None, C.byref(n_maps))
        ids = np.zeros(n_maps.value, dtype=np.int32)
        off = np.zeros(n_maps.value + 1, dtype=np.int64)
        pos = np.zeros(max(tot, 1), dtype=np.float32)
        self.lib.ref_parse_cmap(str(path).encode(), C.c_int(n_maps.value), C.c_int64(tot), _p(ids), _p(off), _p(pos),
                                C.byref(n_maps))
        return {int(ids[k]): pos[off[k]:off[k + 1]].copy() for k in range(n_maps.value)}

    def score_f(self, b, x, i, j, p, q, probs):
        b = np.ascontiguousarray(b, dtype=np.float32)
        x = np.ascontiguousarray(x, dtype=np.float32)
        probs = np.ascontiguousarray(probs, dtype=np.float32)
        return self.lib.ref_score_f(_p(b), b.size, _p(x), x.size, i, j, p, q, _p(probs), probs.size)

    def dp(self, b, x, bprob, xprob, init_mode=0, fitting=0, lookback=6, used=None, swap=0, start_mode=0):
        b = np.ascontiguousarray(b, dtype=np.float32)
        x = np.ascontiguousarray(x, dtype=np.float32)
        bprob = np.ascontiguousarray(bprob, dtype=np.float32)
        xprob = np.ascontiguousarray(xprob, dtype=np.float32)
        nb, nx = b.size - 1, x.size - 1
        S = np.zeros((nb, nx), dtype=np.float32)
        prev = np.zeros((nb, nx, 2), dtype=np.int32)
        u = np.array(sorted(used) if used is not None else [], dtype=np.int32)
        st = np.zeros(2, dtype=np.int32)
        cap = min(nb, nx) + 1
        pjq = np.zeros((cap, 2), dtype=np.int32)
        ps = np.zeros(cap, dtype=np.float32)
        n = self.lib.ref_dp_align(_p(b), C.c_int(b.size), _p(x), C.c_int(x.size), C.c_int(init_mode),
                                  C.c_int(fitting), C.c_int(lookback), _p(u), C.c_int(u.size), _p(xprob), _p(bprob),
                                  C.c_int(swap), C.c_int(start_mode), _p(S), _p(prev), _p(st), _p(pjq), _p(ps))
        return S, prev, (int(st[0]), int(st[1])), (pjq[:n, 0].copy(), pjq[:n, 1].copy(), ps[:n].copy())

    def multi_backtrack(self, b, x, bprob, xprob, n, lookback=6, swap=1):
        b = np.ascontiguousarray(b, dtype=np.float32)
        x = np.ascontiguousarray(x, dtype=np.float32)
        bprob = np.ascontiguousarray(bprob, dtype=np.float32)
        xprob = np.ascontiguousarray(xprob, dtype=np.float32)
        out = np.zeros(max(n, 1), dtype=np.float32)
        got = self.lib.ref_multi_backtrack(_p(b), C.c_int(b.size), _p(x), C.c_int(x.size), C.c_int(lookback),
                                           _p(xprob), _p(bprob), C.c_int(swap), C.c_int(n), _p(out))
        return out[:got].copy()

    def thresholds(self, seg_of, contig_of, score, segs: dict, contigs: dict, pvalue, n_detect=500):
        def flat(maps):
            ids = np.array(sorted(maps), dtype=np.int32)
            off = np.zeros(ids.size + 1, dtype=np.int64)
            for k, i in enumerate(ids):
                off[k + 1] = off[k] + len(maps[i])
            pos = np.concatenate([np.asarray(maps[i], dtype=np.float32) for i in ids])
            return ids, off, pos
        si, so, sp = flat(segs)
        ci, co, cp = flat(contigs)
        seg_of = np.ascontiguousarray(seg_of, dtype=np.int32)
        contig_of = np.ascontiguousarray(contig_of, dtype=np.int32)
        score = np.ascontiguousarray(score, dtype=np.float32)
        out = np.zeros(si.size, dtype=np.float32)
        self.lib.ref_compute_score_thresholds(C.c_int64(score.size), _p(seg_of), _p(contig_of), _p(score),
                                              C.c_int(si.size), _p(si), _p(so), _p(sp), C.c_int(ci.size), _p(ci),
                                              _p(co), _p(cp), C.c_double(pvalue), C.c_int(n_detect), _p(out))
        return {int(si[k]): out[k] for k in range(si.size)}


def random_map(rng, n, integer=False, scale=8000.0):
    iv = rng.lognormal(np.log(scale) - 0.9, 1.2, size=n + 1)
    iv = np.maximum(iv, 30.0)
    pos = np.cumsum(iv)
    pos = np.round(pos) if integer else np.round(pos, 1)
    return pos.astype(np.float32)  # n labels + length


def related_maps(rng, nb, nx, integer=False):
    """A contig and a segment that share a stretch, so good alignments (and ties, when integer) exist."""
    b = random_map(rng, nb, integer)
    s = int(rng.integers(0, max(nb - nx, 1)))
    n = min(nx, nb - s)
    core = b[s:s + n] - b[s] + 500.0
    if not integer:
        core = core * rng.normal(1.0, 0.005) + rng.normal(0, 150.0, size=n)
        core = np.sort(core)
    tail = core[-1] + 1000.0
    x = np.concatenate([core, [tail]])
    if n < nx:
        extra = random_map(rng, nx - n, integer)
        x = np.concatenate([core, core[-1] + extra])
    x = np.round(x) if integer else np.round(x, 1)
    for k in range(1, x.size):
        if x[k] <= x[k - 1]:
            x[k] = x[k - 1] + (1.0 if integer else 0.1)
    return b, x.astype(np.float32)
