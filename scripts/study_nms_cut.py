"""Design study for the NMS kernel ("What to try next on the NMS kernel"): how many warp-rows would still
enter the candidate-emit path if a CTA filtered candidates with a running lower bound of its final top-k cut.  CPU only.

For one synthetic 512 x 512 heatmap per joint (the BASELINE workload) the 5 x 5 positive local maxima are processed in the
kernel's order (64-row strips, rows top to bottom, 4 warps of 128 columns); a candidate enters the emit path iff its score is
>= the cut known when its row is processed.  Cuts compared:
  none        today's kernel (every positive maximum is emitted, the flush keeps the strip's top-k)
  exact       the strip's exact k-th largest score so far (upper bound of what any scheme can achieve per strip)
  per-warp    min over the 4 warps of each warp's ceil(k/4)-th largest so far (no shared structure; valid because
              4 x ceil(k/4) >= k candidates lie at or above it)
  exact, map  the exact running cut if ONE CTA walked the whole map (8 strips in sequence)
"""
import heapq
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pgmp_b200.synthetic as synthetic  # noqa: E402

K, R, STRIP = 30, 2, 64


def local_maxima(m):
    H, W = m.shape
    pad = np.zeros((H + 2 * R, W + 2 * R), m.dtype)
    pad[R:-R, R:-R] = m
    mx = np.zeros_like(m)
    for dy in range(2 * R + 1):
        for dx in range(2 * R + 1):
            np.maximum(mx, pad[dy:dy + H, dx:dx + W], out=mx)
    return (m > 0) & (m == mx)


class TopK:
    def __init__(self, k):
        self.k, self.h = k, []

    def push(self, v):
        if len(self.h) < self.k:
            heapq.heappush(self.h, v)
        elif v > self.h[0]:
            heapq.heapreplace(self.h, v)

    def cut(self):
        return self.h[0] if len(self.h) == self.k else 0.0


def study(m, strip_rows):
    H, W = m.shape
    is_max = local_maxima(m)
    rows_total = {k: 0 for k in ("none", "exact", "per-warp")}
    emitted = {k: 0 for k in rows_total}
    warp_rows = 0
    for y0 in range(0, H, strip_rows):
        exact, warps = TopK(K), [TopK((K + 3) // 4) for _ in range(4)]
        for y in range(y0, min(y0 + strip_rows, H)):
            cut_exact = exact.cut()
            cut_warp = min(w.cut() for w in warps)
            for w in range(4):
                xs = np.flatnonzero(is_max[y, 128 * w:128 * (w + 1)]) + 128 * w
                warp_rows += 1
                vals = m[y, xs]
                for name, cut in (("none", 0.0), ("exact", cut_exact), ("per-warp", cut_warp)):
                    n = int((vals >= cut).sum())
                    emitted[name] += n
                    rows_total[name] += n > 0
                for v in vals:
                    if v >= cut_exact:
                        exact.push(float(v))
                    if v >= cut_warp:
                        warps[w].push(float(v))
    return warp_rows, rows_total, emitted, int(is_max.sum())


def main():
    sm = synthetic.synth_scoremap(0, 17, 512, 30)
    agg = {}
    for j in range(0, 17, 4):
        for label, rows in (("strip", STRIP), ("map", 512)):
            wr, rt, em, nmax = study(sm[j], rows)
            for k in rt:
                a = agg.setdefault((label, k), [0, 0, 0, 0])
                a[0] += wr
                a[1] += rt[k]
                a[2] += em[k]
                a[3] += nmax
    for (label, k), (wr, rt, em, nmax) in sorted(agg.items()):
        print("%-5s %-9s warp-rows entering the emit path %5.1f %%   candidates emitted %5.1f %% of %d maxima" %
              (label, k, 100.0 * rt / wr, 100.0 * em / nmax, nmax))


if __name__ == "__main__":
    main()
