"""A few decoder steps of the bench workload and nothing else - the command ncu wraps (scripts/gpu_profile.sh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np   # noqa: E402
import torch         # noqa: E402

import bench         # noqa: E402
from wavelet_monodepth_b200.kitti_decoders import SparseDepthWaveProgressiveDecoder   # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
wl_name = sys.argv[2] if len(sys.argv) > 2 else bench.MAIN
wl = bench.WORKLOADS[wl_name]
dec = SparseDepthWaveProgressiveDecoder(np.array(wl["ch"]))
bench.synth_params(dec)
dec = dec.cuda().eval()
feats = [f.cuda() for f in bench.synth_features(wl, wl["per_gpu_batch"], 0, pin=False)]
for _ in range(steps):
    out = dec(feats, bench.THRESH)
torch.cuda.synchronize()
print("done", out["total_ops"])
