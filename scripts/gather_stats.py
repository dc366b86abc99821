"""CPU analysis for the next round's gather design: on the bench workload's masks (oracle run, frame 0),
for every 256-row output tile of the sparse 3x3 layers and every tap row dy, how many DISTINCT source rows do the three
taps (dy,-1), (dy,0), (dy,+1) touch, and in how many runs of consecutive row indices do they come?

Today the kernel gathers 3 x 256 rows per (tile, dy, channel chunk) with 192 four-row loads.  `union` rows is what a
shared stage would fetch; `runs` is the number of 2-D tiled loads that would fetch them as contiguous blocks."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np   # noqa: E402
import torch         # noqa: E402

import bench         # noqa: E402
from oracle import kitti as okitti   # noqa: E402  (analysis script, not product code)

wl_name = sys.argv[1] if len(sys.argv) > 1 else bench.MAIN
wl, sd = bench._cpu_setup(wl_name)
feats = bench.synth_features(wl, 1, 0, pin=False)
torch.set_num_threads(8)
with torch.no_grad():
    out = okitti.sparse_forward(sd, feats, bench.THRESH)


def reflect(q, n):
    q = np.abs(q)
    return np.where(q >= n, 2 * (n - 1) - q, q)


def analyse(name, out_mask, src_index, shift):
    """out_mask (H,W) bool: output pixels; src_index (Hs,Ws) int: row of the source at a pixel or -1; shift: source grid = H>>shift."""
    H, W = out_mask.shape
    ys, xs = np.nonzero(out_mask)
    m = len(ys)
    stats = []
    for t0 in range(0, m, 256):
        y, x = ys[t0:t0 + 256], xs[t0:t0 + 256]
        for dy in (-1, 0, 1):
            rows = []
            for dx in (-1, 0, 1):
                qy, qx = reflect(y + dy, H), reflect(x + dx, W)
                r = src_index[qy >> shift, qx >> shift]
                rows.append(r[r >= 0])
            u = np.unique(np.concatenate(rows))
            runs = 1 + int((np.diff(u) > 1).sum()) if len(u) else 0
            centre = np.unique(rows[1])
            hull = int(u[-1] - u[0] + 1) if len(u) else 0
            loads = {}
            if len(u):
                starts = np.concatenate(([0], np.nonzero(np.diff(u) > 1)[0] + 1))
                lens = np.diff(np.concatenate((starts, [len(u)])))
                for br in (8, 16, 32):
                    loads[br] = int(np.ceil(lens / br).sum())
            stats.append((sum(len(r) for r in rows), len(u), runs, len(u) - len(centre), hull, loads.get(8, 0), loads.get(16, 0),
                          loads.get(32, 0)))
    s = np.array(stats, dtype=np.float64)
    if len(s) == 0:
        print("%-14s no active rows" % name)
        return
    print("%-14s rows %7d tiles %4d | per (tile, dy): gathered today %5.0f, distinct %5.0f (%.2fx fewer), runs %5.1f, "
          "rows beyond the centre tap %5.1f (max %d)" % (name, m, -(-m // 256), s[:, 0].mean(), s[:, 1].mean(),
                                                        s[:, 0].mean() / max(s[:, 1].mean(), 1), s[:, 2].mean(), s[:, 3].mean(), s[:, 3].max()))
    print("%-14s   index hull of the distinct rows: mean %6.0f max %6.0f | tiled loads to fetch the runs with 8/16/32-row boxes: "
          "%.1f / %.1f / %.1f (today: %d gather4 loads)" % ("", s[:, 4].mean(), s[:, 4].max(), s[:, 5].mean(), s[:, 6].mean(),
                                                           s[:, 7].mean(), 192))


for i in (3, 2, 1):
    s = i - 1
    S1 = out[("lowres_mask", s)][0, 0].numpy().astype(bool)
    S2 = out[("upconv0_mask", s)][0, 0].numpy().astype(bool)
    S3 = out[("upsample_mask", s)][0, 0].numpy().astype(bool)
    S4 = out[("upconv1_mask", s)][0, 0].numpy().astype(bool)
    h, w = S2.shape
    # upconv(i,0): outputs on S2 (low-res grid); source = previous level's rows under S1 (dense index stands in for the compact one:
    # consecutive active pixels of a row are consecutive either way)
    idx_lo = np.where(S1, np.cumsum(S1.reshape(-1)).reshape(h, w) - 1, -1)
    analyse("upconv(%d,0)" % i, S2, idx_lo, 0)
    # upconv(i,1): outputs on S4 (hi-res); source 0 = upconv(i,0) rows on S2 at (y>>1, x>>1); source 1 = skip rows (dense) under S3
    idx_s2 = np.where(S2, np.cumsum(S2.reshape(-1)).reshape(h, w) - 1, -1)
    analyse("upconv(%d,1) x0" % i, S4, idx_s2, 1)
    H, W = S4.shape
    idx_skip = np.where(S3, np.arange(H * W).reshape(H, W), -1)
    analyse("upconv(%d,1) x1" % i, S4, idx_skip, 0)
