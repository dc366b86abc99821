"""Micro-benchmark of the NCHW->rows move: plain vs gated at several mask densities (f0/f1 of the bench workload)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch   # noqa: E402

from wavelet_monodepth_b200 import ops   # noqa: E402

dev = "cuda"


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for (n, c, h, w) in ((32, 64, 160, 512), (32, 256, 80, 256)):
    x = torch.rand(n, c, h, w, device=dev)
    us = timed(lambda: ops.nchw_to_rows(x))
    by = 8.0 * x.numel()
    print("plain  %s: %.1f us  %.0f GB/s" % ((n, c, h, w), us, by / us / 1e3), flush=True)
    for cell, p in ((16, 1.0), (16, 0.5), (16, 0.25), (16, 0.1), (1, 0.25), (16, 0.0)):
        g = (torch.rand(n, 1, -(-h // cell), -(-w // cell), device=dev) < p)
        gate = g.repeat_interleave(cell, 2).repeat_interleave(cell, 3)[:, :, :h, :w].contiguous().to(torch.uint8)
        grp = float(gate.reshape(n, -1, 32).any(-1).float().mean())
        us = timed(lambda: ops.nchw_to_rows(x, gate=gate))
        print("gated  cell %2d p %.2f (px %.3f, groups %.3f): %.1f us  eff %.0f GB/s of moved bytes" %
              (cell, p, float(gate.float().mean()), grp, us, by * float(gate.float().mean()) / us / 1e3), flush=True)
