"""A/B of decoder options on the bench workload, CUDA-graph replay, one process:  python scripts/ab_flags.py [steps]"""
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np   # noqa: E402
import torch         # noqa: E402

import bench         # noqa: E402
from wavelet_monodepth_b200 import graphs   # noqa: E402
from wavelet_monodepth_b200.kitti_decoders import SparseDepthWaveProgressiveDecoder   # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
wl = bench.WORKLOADS[bench.MAIN]
dev = torch.device("cuda", 0)
dec = SparseDepthWaveProgressiveDecoder(np.array(wl["ch"]))
bench.synth_params(dec)
dec = dec.to(dev).eval()
resident = [f.to(dev) for f in bench.synth_features(wl, wl["per_gpu_batch"], 0, pin=False)]
ref = None
for ll, oc in itertools.product((0, 1), (0, 1)):
    dec.factored_ll, dec.overlap_compaction = bool(ll), bool(oc)
    g = graphs.GraphedSparseDecoder(dec, resident, bench.THRESH)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    d0 = out[("disp", 0)].clone()
    if ref is None:
        ref = d0
    print(json.dumps({"factored_ll": ll, "overlap_compaction": oc, "graph_ms": round(ms, 4), "fps": round(32e3 / ms, 1),
                      "launches": g.launches, "max_rel_diff_disp0": float((d0 - ref).abs().max() / ref.abs().max()),
                      "total_ops": out["total_ops"]}), flush=True)
    del g
