"""Timeline of one eager decoder step on the bench workload: every libwmd launch with its start / end offset (CUDA events
on its own stream, relative to the step's first event) and stream - shows what the tensor-bound convolutions wait for.

    python scripts/step_timeline.py [channels_last]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from wavelet_monodepth_b200 import ops
from wavelet_monodepth_b200.kitti_decoders import SparseDepthWaveProgressiveDecoder

wl = bench.WORKLOADS[bench.MAIN]
dec = SparseDepthWaveProgressiveDecoder(np.array(wl["ch"])); bench.synth_params(dec); dec = dec.cuda().eval()
feats = [f.cuda() for f in bench.synth_features(wl, wl["per_gpu_batch"], 0, pin=False)]
if len(sys.argv) > 1 and sys.argv[1] == "channels_last":
    feats = [f.contiguous(memory_format=torch.channels_last) for f in feats]
for _ in range(3):
    dec(feats, bench.THRESH)
torch.cuda.synchronize()
prof = ops.Profiler(); ops.set_profiler(prof)
origin = torch.cuda.Event(enable_timing=True); origin.record()
dec(feats, bench.THRESH)
fin = torch.cuda.Event(enable_timing=True); fin.record()
torch.cuda.synchronize(); ops.set_profiler(None)
streams = {}
rows = []
for name, s, e, info in prof.records:
    sid = streams.setdefault(info.get("_stream"), len(streams))
    extra = ""
    if name.startswith("conv_rows"):
        extra = "%dx [%d,%d]->%d" % (info["taps"], info["c0"], info["c1"], info["cout"])
    rows.append((origin.elapsed_time(s) * 1e3, origin.elapsed_time(e) * 1e3, sid, name, extra))
rows.sort()
print("step %.0f us (eager, with profiling events)" % (origin.elapsed_time(fin) * 1e3))
busy_end = 0.0
for st, en, sid, name, extra in rows:
    gap = st - busy_end
    print("%8.0f %8.0f  %6.0f us  s%d %s%-18s %s%s" % (st, en, en - st, sid, "    " * sid, name, extra,
                                                    ("   <- %.0f us after everything earlier ended" % gap) if gap > 5 else ""))
    busy_end = max(busy_end, en)
