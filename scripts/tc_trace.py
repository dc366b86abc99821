"""Phase timing of the tcgen05 conv kernel: builds libwmd_trace.so (-DWMD_TC_TRACE), runs one layer and prints the
mean clocks between the trace points of CTA 0 (split warp 0, gather warp 8, issuer 0).

    python scripts/tc_trace.py build          # here (cross-compile)
    python scripts/tc_trace.py run [layer]    # on the GPU box
"""
import ctypes
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from wavelet_monodepth_b200 import build as wbuild   # noqa: E402

TRACE_LIB = os.path.join(REPO, "scripts", "bench_cu", "_bin", "libwmd_trace%s.so" % os.environ.get("WMD_TRACE_TAG", ""))

if sys.argv[1] == "build":
    os.makedirs(os.path.dirname(TRACE_LIB), exist_ok=True)
    cmd = [wbuild.nvcc_path(), "-DWMD_TC_TRACE"] + os.environ.get("WMD_TRACE_FLAGS", "").split() + wbuild.NVCC_FLAGS + ["-I", os.path.join(REPO, "include"), "-I", wbuild.CSRC,
                                                                       "-o", TRACE_LIB] + wbuild.sources()
    subprocess.run(cmd, check=True)
    print(TRACE_LIB)
    sys.exit(0)

import numpy as np   # noqa: E402
import torch   # noqa: E402
from wavelet_monodepth_b200 import _lib   # noqa: E402
_lib.LIB_PATH = TRACE_LIB
from wavelet_monodepth_b200 import ops   # noqa: E402
from wavelet_monodepth_b200._lib import ACT_ELU, PAD_REFLECT   # noqa: E402

LAYERS = {"upconv41": (20, 64, 256, 1024, 256, 9), "upconv31": (40, 128, 128, 512, 128, 9), "upconv21": (80, 256, 64, 256, 64, 9),
          "upconv11": (160, 512, 32, 64, 32, 9), "upconv40": (10, 32, 2048, 0, 256, 9)}
name = sys.argv[2] if len(sys.argv) > 2 else "upconv41"
splits = int(sys.argv[3]) if len(sys.argv) > 3 else 1
h, w, c0, c1, cout, taps = LAYERS[name]
n = 32
dev = "cuda"
torch.manual_seed(0)
x0 = torch.rand(n * (h // 2) * (w // 2) if c1 else n * h * w, c0, device=dev)
x1 = torch.rand(n * h * w, c1, device=dev) if c1 else None
wt = (torch.rand(cout, c0 + c1, 3, 3, device=dev) - 0.5) * 0.1
wp = ops.pack_weight(wt, c1, kind="tc")
for _ in range(2):
    ops.conv_rows(x0, c0, wp, None, cout, n, h, w, taps=taps, pad=PAD_REFLECT, act=ACT_ELU, shift0=1 if c1 else 0, x1=x1, c1=c1,
                  splits=splits)
torch.cuda.synchronize()
lib = _lib.load()
K, S = 256, 8
buf = np.zeros(3 * S * K, dtype=np.int64)
assert lib.wmd_debug_tc_trace(buf.ctypes.data_as(ctypes.c_void_p)) == 0
t = buf.reshape(3, S, K)
nch = taps * (-(-c0 // 32) + -(-c1 // 32))
lo, hi = 40, min(K, nch) - 8                       # steady state of the first tile (skips the epoch boundaries' neighbours)
sel = [c for c in range(lo, hi) if c % 32 not in (0, 1, 31)]
print("%s: %d chunks/tile; clocks per chunk (mean over %d steady chunks)" % (name, nch, len(sel)))
for role, nm, labels in ((0, "split warp 0", ["wait raw A", "wait MMA(c-2)", "LDS+split+tcgen05.st", "wait::st + arrive", "loop"]),
                         (1, "gather warp 8", ["wait raw stage free", "issue gather c+2", "wait gather c + arrive", "loop"]),
                         (2, "issuer 0", ["wait weight image (B)", "wait split A", "issue MMAs + commit", "loop"])):
    tt = t[role]
    npts = len(labels)
    period = np.mean([tt[0, c + 1] - tt[0, c] for c in sel])
    print("  %-16s period %7.0f clk" % (nm, period))
    for i in range(npts - 1):
        print("      %-24s %7.0f" % (labels[i], np.mean([tt[i + 1, c] - tt[i, c] for c in sel])))
    print("      %-24s %7.0f" % ("-> next chunk top", np.mean([tt[0, c + 1] - tt[npts - 1, c] for c in sel])))
