"""GB/s of the layout moves on the bench shapes (A/B of tile shapes: WMD_LIB_PATH selects the build)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wavelet_monodepth_b200 import _lib
if os.environ.get("WMD_LIB_PATH"):
    _lib.LIB_PATH = os.path.abspath(os.environ["WMD_LIB_PATH"])
from wavelet_monodepth_b200 import ops


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


def blob_mask(n, h, w, density, cell=8, seed=0):
    g = torch.Generator().manual_seed(seed)
    m = (torch.rand(n, 1, h // cell, w // cell, generator=g) < density).to(torch.uint8)
    return m.repeat_interleave(cell, 2).repeat_interleave(cell, 3).contiguous().cuda()


out = []
for name, (n, c, h, w) in (("f4", (32, 2048, 10, 32)), ("skip4", (32, 1024, 20, 64))):
    x = torch.rand(n, c, h, w, device="cuda")
    us = timed(lambda: ops.nchw_to_rows(x))
    out.append("%s nchw_to_rows %.0f us %.0f GB/s" % (name, us, 2 * x.numel() * 4 / us / 1e3))
n, c, h, w = 32, 512, 40, 128
x = torch.rand(n, c, h, w, device="cuda"); gate = blob_mask(n, h, w, 0.35)
us = timed(lambda: ops.nchw_to_rows(x, gate=gate))
out.append("skip3 gated(%.2f) %.0f us %.0f GB/s(marked)" % (float(gate.float().mean()), us, 2 * float(gate.sum()) * c * 4 / us / 1e3))
for name, (n, c, h, w, dens) in (("skip2", (32, 256, 80, 256, 0.19)), ("skip1", (32, 64, 160, 512, 0.10))):
    x = torch.rand(n, c, h, w, device="cuda"); mask = blob_mask(n, h, w, dens, seed=1)
    _, pixels, offsets = ops.compact(mask, want_idxmap=False)
    m = int(offsets[n])
    us = timed(lambda: ops.gather_rows_list(x, pixels, offsets[n:]))
    out.append("%s gather_rows_list(%d rows) %.0f us %.0f GB/s" % (name, m, us, 2 * m * c * 4 / us / 1e3))
print(os.environ.get("WMD_LIB_PATH", "default"), " | ".join(out))
