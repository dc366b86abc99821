"""Quick bounded check of the shared three-tap gather: tc (share on) vs the fp32 FMA engine on dense, clustered-sparse and
scattered-sparse (overflow fallback) 3x3 layers, with and without the second (skip) source."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch   # noqa: E402

from wavelet_monodepth_b200 import ops   # noqa: E402
from wavelet_monodepth_b200._lib import ACT_ELU, PAD_REFLECT, PAD_ZERO   # noqa: E402

dev = "cuda"
torch.manual_seed(0)
ok = True
for name, n, h, w, c0, c1, cout, mode, pad in (
        ("dense K45", 2, 24, 64, 160, 0, 64, "dense", PAD_REFLECT),
        ("dense skip", 2, 24, 64, 64, 96, 128, "dense", PAD_REFLECT),
        ("cluster", 4, 40, 128, 128, 0, 64, "cluster", PAD_REFLECT),
        ("cluster skip K54", 4, 40, 128, 64, 128, 32, "cluster", PAD_ZERO),
        ("scatter", 4, 40, 128, 96, 0, 64, "scatter", PAD_REFLECT),
        ("scatter skip", 2, 40, 128, 32, 64, 32, "scatter", PAD_REFLECT)):
    total = n * h * w
    if mode == "dense":
        pixels = count = None
        m = total
    else:
        if mode == "cluster":
            g = torch.rand(n, 1, h // 8, w // 8, device=dev) < 0.4
            mask = g.repeat_interleave(8, 2).repeat_interleave(8, 3)
        else:
            mask = torch.rand(n, 1, h, w, device=dev) < 0.3
        _, pixels, off = ops.compact(mask.to(torch.uint8))
        count = off[n:]
        m = int(count.item())
    if c1:
        x0 = torch.rand(n * (h // 2) * (w // 2), c0, device=dev)
        x1 = torch.rand(total, c1, device=dev)
    else:
        x0, x1 = torch.rand(total, c0, device=dev), None
    wt = (torch.rand(cout, c0 + c1, 3, 3, device=dev) - 0.5) * 0.1
    bias = torch.rand(cout, device=dev)
    ys = {}
    for kind in ("simt", "tc"):
        wp = ops.pack_weight(wt, c1, kind=kind)
        kw = dict(taps=9, pad=pad, act=ACT_ELU, shift0=1 if c1 else 0, x1=x1, c1=c1)
        if pixels is not None:
            kw.update(pixels=pixels, count=count)
        ys[kind] = ops.conv_rows(x0, c0, wp, bias, cout, n, h, w, **kw)[:m].clone()
        torch.cuda.synchronize()
    err = float((ys["tc"] - ys["simt"]).abs().max() / ys["simt"].abs().max())
    print("%-18s rows %7d  rel diff tc vs simt %.2e %s" % (name, m, err, "ok" if err < 1e-5 else "FAIL"), flush=True)
    ok = ok and err < 1e-5
print("ALL OK" if ok else "FAILED")
