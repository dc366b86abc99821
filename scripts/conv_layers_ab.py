"""A/B of the gather-GEMM launches of one bench step: per-layer CUDA-event times with the shared-tap gather off / on.
    python scripts/conv_layers_ab.py            (GPU box)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np   # noqa: E402
import torch         # noqa: E402

import bench         # noqa: E402
from wavelet_monodepth_b200 import _lib, ops   # noqa: E402
from wavelet_monodepth_b200.kitti_decoders import SparseDepthWaveProgressiveDecoder   # noqa: E402

wl = bench.WORKLOADS[bench.MAIN]
dec = SparseDepthWaveProgressiveDecoder(np.array(wl["ch"]))
bench.synth_params(dec)
dec = dec.cuda().eval()
feats = [f.cuda() for f in bench.synth_features(wl, wl["per_gpu_batch"], 0, pin=False)]
lib = _lib.load()
tables = {}
for sh in (0, 1):
    lib.wmd_conv_tc_set_shared_taps(sh)
    dec(feats, bench.THRESH)
    prof = ops.Profiler()
    torch.cuda.synchronize()
    ops.set_profiler(prof)
    steps = 3
    for _ in range(steps):
        dec(feats, bench.THRESH)
    torch.cuda.synchronize()
    ops.set_profiler(None)
    tables[sh] = bench.conv_layer_table(prof.results(), 6570.9, 761.6, steps)
tot = [0.0, 0.0]
for a, b in zip(tables[0], tables[1]):
    tot[0] += a["us"]; tot[1] += b["us"]
    print("taps %d cin %-12s cout %3d rows %6d : per-tap %7.1f us   shared %7.1f us   (%.2fx)  exec frac %.3f -> %.3f" % (
        a["taps"], a["cin"], a["cout"], a["rows"], a["us"], b["us"], a["us"] / b["us"], a.get("tensor_frac_executed", 0), b.get("tensor_frac_executed", 0)))
print("sum %.1f -> %.1f us" % tuple(tot))
