import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from wavelet_monodepth_b200.kitti_decoders import SparseDepthWaveProgressiveDecoder
wl = bench.WORKLOADS[bench.MAIN]
dec = SparseDepthWaveProgressiveDecoder(np.array(wl["ch"])); bench.synth_params(dec); dec = dec.cuda().eval()
feats = [f.cuda() for f in bench.synth_features(wl, 32, 0, pin=False)]
def stats():
    s = torch.cuda.memory_stats()
    return s["num_device_alloc"], s["num_device_free"], s["num_alloc_retries"], s["reserved_bytes.all.current"] >> 20, s["allocated_bytes.all.peak"] >> 20
for i in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dec(feats, bench.THRESH)
    torch.cuda.synchronize(); print("step", i, "%.2f ms" % (1e3 * (time.perf_counter() - t0)), stats(), flush=True)
