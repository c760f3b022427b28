"""Developer probe (CPU, numpy + the oracle): how many of a cell's 36 candidates could be dropped WARP-UNIFORMLY by a
bound that costs no MUFU?  For C2-shaped (contig, segment) pairs: S from the oracle, then per cell all 36 candidate scores
(float64 statistics, not the product arithmetic), cheap upper bounds, and per 32-row strip the fraction of look-back rows /
single candidates whose cheap bound is below a threshold every lane knows after row dj = 1."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import helpers  # noqa: E402
from ampliconreconstructorom_b200 import synth  # noqa: E402

LB = 6


def study(b, x, xprob, S, bound):
    nb, nx = S.shape
    NEG = -np.inf
    Sp = np.full((nb + LB, nx + LB), NEG)
    Sp[LB:, LB:] = S
    bb = np.concatenate([np.full(LB, np.nan), b[:nb]])
    xx = np.concatenate([np.full(LB, np.nan), x[:nx]])
    cp = np.concatenate([[0.0], np.cumsum(xprob.astype(np.float64))])  # cp[k] = sum prob[0..k-1]
    exact = np.full((LB, LB, nb, nx), NEG)
    hi0 = np.full((LB, LB, nb, nx), NEG)
    for dj in range(1, LB + 1):
        db = bb[LB:] - bb[LB - dj:LB - dj + nb]
        for dq in range(1, LB + 1):
            dxv = xx[LB:] - xx[LB - dq:LB - dq + nx]
            q = np.arange(nx)
            p = q - dq
            fp = np.where(p >= 0, 5000.0 * (cp[np.maximum(q, 0)] - cp[np.maximum(p + 1, 0)]), 0.0)
            sc = Sp[LB - dj:LB - dj + nb, LB - dq:LB - dq + nx]
            d = np.abs(db[:, None] - dxv[None, :])
            fnfp = 5000.0 * (dj - 1) + fp[None, :]
            with np.errstate(invalid="ignore"):
                exact[dj - 1, dq - 1] = np.where(np.isfinite(sc) & np.isfinite(d), sc + 10000.0 - fnfp - d ** 1.2, NEG)
                hi0[dj - 1, dq - 1] = np.where(np.isfinite(sc) & np.isfinite(d), sc + 10000.0 - fnfp - bound(d), NEG)
    best = exact.reshape(36, nb, nx).max(0)
    row1 = exact[0].max(0)  # what a lane knows (to within the tight bound's slack) after row dj = 1
    out = {}
    nstrips = (nb + 31) // 32
    pad = nstrips * 32 - nb

    def strips(a, fill):  # (nb, nx) -> (nstrips, 32, nx)
        return np.concatenate([a, np.full((pad, nx), fill)]).reshape(nstrips, 32, nx)

    live = strips(np.isfinite(best), False)
    steps = live.any(1)
    out["steps"] = int(steps.sum())
    # per-lane: candidate needed iff hi0 >= threshold (row-1 exact best)
    for name, thr in (("thr=row1", row1), ("thr=best", best)):
        need = hi0 >= thr[None, None]  # (6,6,nb,nx)
        need_row = need.any(1)  # (6, nb, nx)
        lane_rows = [need_row[dj][np.isfinite(best)].mean() for dj in range(LB)]
        warp_rows = []
        for dj in range(LB):
            w = strips(need_row[dj], False).any(1)
            warp_rows.append(w[steps].mean())
        warp_cand = np.zeros((LB, LB))
        for dj in range(LB):
            for dq in range(LB):
                w = strips(need[dj, dq], False).any(1)
                warp_cand[dj, dq] = w[steps].mean()
        out[name] = (lane_rows, warp_rows, warp_cand)
    win = exact.reshape(36, nb, nx).argmax(0)[np.isfinite(best)]
    out["winner_row_hist"] = np.bincount(win // LB, minlength=LB) / max(win.size, 1)
    return out


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    npairs = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    segs, contigs = synth.make_workload(seed, 40, 12)
    orc = helpers.Oracle()
    cp_, sp_ = orc.probs(contigs, 2), orc.probs(segs, 2)
    rng = np.random.default_rng(1)
    bounds = {
        "d>=1?d:0": lambda d: np.where(d >= 1.0, d, 0.0),
        "2 tangents(1k,8k)": lambda d: np.maximum(np.maximum(1.2 * 1000 ** 0.2 * d - 0.2 * 1000 ** 1.2,
                                                             1.2 * 8000 ** 0.2 * d - 0.2 * 8000 ** 1.2), 0.0),
        "int-trick 0.877..1": lambda d: 0.877 * d ** 1.2,
        "exact (ceiling)": lambda d: d ** 1.2,
    }
    for bname, bound in bounds.items():
        acc = None
        tot = 0
        for k in range(npairs):
            c = int(rng.integers(1, len(contigs) + 1))
            s = int(rng.integers(1, len(segs) + 1))
            b = np.asarray(contigs[c], dtype=np.float32)
            x = np.asarray(segs[s], dtype=np.float32)
            S, _ = orc.fill(b, x, cp_[c], sp_[s], 0, 0, 6, None, 0)[:2]
            r = study(b.astype(np.float64), x.astype(np.float64), sp_[s], S.astype(np.float64), bound)
            wgt = r["steps"]
            tot += wgt
            vals = [np.array(r["thr=row1"][0]), np.array(r["thr=row1"][1]), r["thr=row1"][2],
                    np.array(r["thr=best"][1]), r["thr=best"][2], r["winner_row_hist"]]
            acc = [v * wgt for v in vals] if acc is None else [a + v * wgt for a, v in zip(acc, vals)]
        acc = [a / tot for a in acc]
        np.set_printoptions(precision=3, suppress=True, linewidth=150)
        print(f"== bound {bname}  ({npairs} pairs, {tot} warp-steps)")
        print(" per-lane   P(row dj needed | thr=row-1 best):", acc[0])
        print(" per-WARP   P(row dj needed | thr=row-1 best):", acc[1], " mean rows 2..6 needed:", acc[1][1:].sum())
        print(" per-WARP   P(row dj needed | thr=final best):", acc[3], " mean rows 2..6 needed:", acc[3][1:].sum())
        print(" per-WARP candidates needed (thr=row1): total", acc[2].sum(), "of 36")
        print(acc[2])
        print(" per-WARP candidates needed (thr=best): total", acc[4].sum(), "of 36")
        print(" winner row histogram:", acc[5])


if __name__ == "__main__":
    main()
