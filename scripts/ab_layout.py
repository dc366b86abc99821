"""A/B of the layout-move options on the bench workload (one process, one set of features):
   device-resident graph replay and the end-to-end step (pinned host features) for every (gated, overlap) pair,
   plus the zero-copy variant of the end-to-end step.  Prints one JSON line per variant.  Run under gpurun."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np   # noqa: E402
import torch         # noqa: E402

import bench         # noqa: E402
from wavelet_monodepth_b200 import graphs   # noqa: E402
from wavelet_monodepth_b200.kitti_decoders import SparseDepthWaveProgressiveDecoder   # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
wl_name = sys.argv[2] if len(sys.argv) > 2 else bench.MAIN
wl = bench.WORKLOADS[wl_name]
n = wl["per_gpu_batch"]
dev = torch.device("cuda", 0)
dec = SparseDepthWaveProgressiveDecoder(np.array(wl["ch"]))
bench.synth_params(dec)
dec = dec.to(dev).eval()
host = bench.synth_features(wl, n, 0, pin=True)
resident = [f.to(dev) for f in host]
disp_host = torch.empty((n, 1, wl["height"], wl["width"]), dtype=torch.float32).pin_memory()
copy_stream = torch.cuda.Stream()


def timed(fn, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def e2e(zero_copy):
    dma = [k for k in range(5) if k not in zero_copy]
    bufs = [[host[k] if k in zero_copy else torch.empty_like(host[k], device=dev) for k in range(5)] for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    gs = [graphs.GraphedSparseDecoder(dec, b, bench.THRESH) for b in bufs]
    state = {"i": 0}

    def enqueue(slot):
        with torch.cuda.stream(copy_stream):
            for k in dma:
                bufs[slot][k].copy_(host[k], non_blocking=True)
            ready[slot].record(copy_stream)

    def step():
        slot = state["i"] % 2
        enqueue(1 - slot)
        torch.cuda.current_stream().wait_event(ready[slot])
        o = gs[slot].replay()
        disp_host.copy_(o[("disp", 0)], non_blocking=True)
        copy_stream.wait_stream(torch.cuda.current_stream())
        state["i"] += 1

    enqueue(0)
    ms = timed(step, warm=2)
    del gs, bufs
    torch.cuda.empty_cache()
    return ms


ref = None
for gated, overlap in ((0, 0), (1, 0), (0, 1), (1, 1)):
    dec.gated_layout, dec.overlap_layout = bool(gated), bool(overlap)
    g = graphs.GraphedSparseDecoder(dec, resident, bench.THRESH)
    ms = timed(g.replay)
    out = g.replay()
    d0 = out[("disp", 0)].clone()
    if ref is None:
        ref = d0
    same = bool(torch.equal(d0, ref))
    del g
    ms_eager = timed(lambda: dec(resident, bench.THRESH))
    line = {"gated": gated, "overlap": overlap, "graph_ms": round(ms, 3), "fps": round(n / ms * 1e3, 1),
            "eager_ms": round(ms_eager, 3), "bit_identical": same, "e2e_dma_ms": round(e2e(()), 3)}
    if gated:
        for zc in ((0,), (0, 1), (0, 1, 2)):
            line["e2e_zero_copy_%s_ms" % "".join(map(str, zc))] = round(e2e(zc), 3)
        up = {k: out[("upsample_mask", k)] for k in (0, 1, 2)}
        line["group_density"] = {k: round(float(v.reshape(n, -1, 32).any(-1).float().mean()), 4) for k, v in up.items()}
        line["pixel_density"] = {k: round(float(v.float().mean()), 4) for k, v in up.items()}
    print(json.dumps(line), flush=True)
