"""Per-tile timeline of the tcgen05 conv kernel (CTA 0) for the short-reduction layers (1x1 head stages).
    python scripts/tc_trace.py build ; then on the GPU box: python scripts/tc_tile_trace.py"""
import ctypes
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np   # noqa: E402
import torch   # noqa: E402
from wavelet_monodepth_b200 import _lib   # noqa: E402
_lib.LIB_PATH = os.path.join(REPO, "scripts", "bench_cu", "_bin", "libwmd_trace.so")
from wavelet_monodepth_b200 import ops   # noqa: E402
from wavelet_monodepth_b200._lib import ACT_LRELU, PAD_REFLECT   # noqa: E402

# name: (rows, c0, cout, taps)
LAYERS = {"t4": (40960, 256, 576, 1), "z4": (40960, 576, 54, 1), "t3": (69892, 128, 256, 1), "z3": (69892, 256, 54, 1),
          "z2": (160314, 128, 54, 1), "up30": (25658, 256, 128, 9)}
dev = "cuda"
lib = _lib.load()
names = sys.argv[1:] or list(LAYERS)
for name in names:
    rows, c0, cout, taps = LAYERS[name]
    n, h, w = 32, 40, 128 * ((rows + 32 * 40 * 128 - 1) // (32 * 40 * 128))
    total = n * h * w
    torch.manual_seed(0)
    x0 = torch.rand(total, c0, device=dev)
    wt = (torch.rand(cout, c0, 3 if taps == 9 else 1, 3 if taps == 9 else 1, device=dev) - 0.5) * 0.1
    wp = ops.pack_weight(wt, 0, kind="tc")
    pixels = torch.arange(rows, device=dev, dtype=torch.int32)
    count = torch.tensor([rows], device=dev, dtype=torch.int32)
    kw = dict(taps=taps, pad=PAD_REFLECT, act=ACT_LRELU, act_param=0.1, pixels=pixels, count=count)
    for _ in range(2):
        ops.conv_rows(x0, c0, wp, None, cout, n, h, w, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.conv_rows(x0, c0, wp, None, cout, n, h, w, **kw)
    e1.record()
    torch.cuda.synchronize()
    buf = np.zeros(64 * 8, dtype=np.int64)
    assert lib.wmd_debug_tc_tile_trace(buf.ctypes.data_as(ctypes.c_void_p)) == 0
    t = buf.reshape(64, 8)
    bn = lib.wmd_conv_tc_tile_n(cout)
    tiles = -(-rows // 256) * -(-cout // bn)
    per_cta = -(-tiles // 148)
    print("%s: rows %d K %d N %d (tile N %d): %.1f us, %d tiles, <= %d per CTA, %d chunks/tile" %
          (name, rows, c0 * taps, cout, bn, e0.elapsed_time(e1) * 1e3, tiles, per_cta, taps * -(-c0 // 32)))
    labels = ["tables", "first raw A", "chunk loop", "last epoch wait", "drain", "store", "end barrier", "-> next tile"]
    k = min(per_cta, 6)
    for i in range(k):
        d = [t[i, j + 1] - t[i, j] for j in range(7)] + [t[i + 1, 0] - t[i, 7] if i + 1 < k else 0]
        print("   tile %d: " % i + "  ".join("%s %d" % (l, v) for l, v in zip(labels, d)) + "   total %d clk" % (t[i, 7] - t[i, 0]))
