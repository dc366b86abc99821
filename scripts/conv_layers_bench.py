"""Per-layer timing of the two gather-GEMM engines on the dense-equivalent conv shapes of a workload."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch   # noqa: E402

from wavelet_monodepth_b200 import ops   # noqa: E402
from wavelet_monodepth_b200._lib import ACT_ELU, PAD_REFLECT   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
kinds = sys.argv[2].split(",") if len(sys.argv) > 2 else ["simt", "tc"]
# (name, h, w, c0, c1(skip, at 2h x 2w grid => shift), cout, taps, density)
LAYERS = [
    ("upconv40", 10, 32, 2048, 0, 256, 9, 1.0),
    ("upconv41", 20, 64, 256, 1024, 256, 9, 1.0),
    ("head4_1x1", 20, 64, 256, 0, 576, 1, 1.0),
    ("upconv30", 20, 64, 256, 0, 128, 9, 0.6),
    ("upconv31", 40, 128, 128, 512, 128, 9, 0.5),
    ("head3_1x1", 40, 128, 128, 0, 256, 1, 0.5),
    ("upconv20", 40, 128, 128, 0, 64, 9, 0.35),
    ("upconv21", 80, 256, 64, 256, 64, 9, 0.3),
    ("head2_1x1", 80, 256, 64, 0, 128, 1, 0.3),
    ("upconv10", 80, 256, 64, 0, 32, 9, 0.2),
    ("upconv11", 160, 512, 32, 64, 32, 9, 0.15),
    ("head1_1x1", 160, 512, 32, 0, 64, 1, 0.15),
]
only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
if only:
    LAYERS = [l for l in LAYERS if l[0] in only]
dev = "cuda"
torch.manual_seed(0)
for name, h, w, c0, c1, cout, taps, dens in LAYERS:
    rows_cap = n * h * w
    m = int(rows_cap * dens)
    count = torch.tensor([m], dtype=torch.int32, device=dev)
    pixels = torch.sort(torch.randperm(rows_cap, device=dev)[:m])[0].to(torch.int32) if dens < 1 else None
    if c1:
        x0 = torch.rand(n * (h // 2) * (w // 2), c0, device=dev)
        x1 = torch.rand(rows_cap, c1, device=dev)
    else:
        x0 = torch.rand(rows_cap, c0, device=dev)
        x1 = None
    wt = (torch.rand(cout, c0 + c1, 3 if taps == 9 else 1, 3 if taps == 9 else 1, device=dev) - 0.5) * 0.1
    bias = torch.rand(cout, device=dev)
    res = {}
    outs = {}
    for kind in kinds:
        wp = ops.pack_weight(wt, c1, kind=kind.split("@")[0])      # "tc@0" / "tc@1" / "tc@3": force the split mode
        kw = dict(taps=taps, pad=PAD_REFLECT, act=ACT_ELU, shift0=1 if c1 else 0, x1=x1, c1=c1)
        if pixels is not None:
            kw.update(pixels=pixels, count=count)
        if "@" in kind:
            kw.update(splits=int(kind.split("@")[1]))
        for _ in range(2):
            y = ops.conv_rows(x0, c0, wp, bias, cout, n, h, w, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            y = ops.conv_rows(x0, c0, wp, bias, cout, n, h, w, **kw)
        e1.record()
        torch.cuda.synchronize()
        res[kind] = e0.elapsed_time(e1) / 5
        outs[kind] = y[:m].clone()
    fl = 2.0 * taps * (c0 + c1) * cout * m
    err = float((outs[kinds[0]] - outs[kinds[-1]]).abs().max() / outs[kinds[0]].abs().max()) if len(kinds) > 1 else 0
    print("%-10s rows %8d K %6d N %4d  " % (name, m, taps * (c0 + c1), cout) +
          "  ".join("%s %8.3f ms %6.1f TF/s" % (k, res[k], fl / res[k] / 1e9) for k in kinds) + "   rel diff %.1e" % err,
          flush=True)
