"""Fused 1x1 head stages (wmd_head_mlp_f32) vs the two gather-GEMM launches they replace, at the bench workload's sizes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch   # noqa: E402

from wavelet_monodepth_b200 import ops   # noqa: E402
from wavelet_monodepth_b200._lib import ACT_LRELU   # noqa: E402

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


for c, rows in ((64, 160314), (32, 350128), (64, 1310720), (32, 2621440)):
    n1, nz = 2 * c, 54
    torch.manual_seed(0)
    x = torch.rand(rows, c, device=dev)
    w1 = (torch.rand(n1, c, 1, 1, device=dev) - 0.5) * 0.3
    b1 = torch.rand(n1, device=dev) * 0.1
    wz = (torch.rand(nz, n1, 1, 1, device=dev) - 0.5) * 0.3
    packed = ops.pack_head_mlp(w1, b1, wz)
    cnt = torch.tensor([rows], dtype=torch.int32, device=dev)
    pix = torch.arange(rows, dtype=torch.int32, device=dev)
    us_f = timed(lambda: ops.head_mlp(x, c, packed, n1, 0.1, count=cnt, max_rows=rows))
    wp1, wpz = ops.pack_weight(w1), ops.pack_weight(wz)

    def two():
        t = ops.conv_rows(x, c, wp1, b1, n1, 1, 1, rows, taps=1, act=ACT_LRELU, act_param=0.1, pixels=pix, count=cnt)
        return ops.conv_rows(t, n1, wpz, None, nz, 1, 1, rows, taps=1, pixels=pix, count=cnt)

    us_2 = timed(two)
    z, z2 = ops.head_mlp(x, c, packed, n1, 0.1, count=cnt, max_rows=rows), two()
    err = float((z[:, :nz] - z2[:, :nz]).abs().max() / z2[:, :nz].abs().max())
    by = 4.0 * rows * (c + 56)
    fl = 2.0 * rows * (c * n1 + n1 * nz)
    print("c %3d rows %8d: fused %7.1f us (%.0f GB/s, %.1f TF/s fp32-eq)   two launches (%s, %s) %7.1f us   rel diff %.1e"
          % (c, rows, us_f, by / us_f / 1e3, fl / us_f / 1e6, wp1.kind, wpz.kind, us_2, err), flush=True)
