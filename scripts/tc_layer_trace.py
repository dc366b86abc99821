"""Phase timing of conv_rows_tc on the REAL bench workload: runs the sparse decoder on the bench's synthetic step with
the tracing build (scripts/tc_trace.py build) and prints, for the k-th gather-GEMM launch of the forward, CTA 0's
per-tile timeline and the mean clocks between the per-chunk trace points of split warp 0 / gather warp 8 / issuer 0.

    python scripts/tc_trace.py build                         # here (cross-compile)
    python scripts/tc_layer_trace.py <k>[,<k>...] [sh=0|1]   # on the GPU box; k = 0..11 in launch order
"""
import ctypes
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np   # noqa: E402
import torch   # noqa: E402
from wavelet_monodepth_b200 import _lib   # noqa: E402
_lib.LIB_PATH = os.path.join(REPO, "scripts", "bench_cu", "_bin", "libwmd_trace.so")
from wavelet_monodepth_b200 import ops, synth   # noqa: E402
from wavelet_monodepth_b200.kitti_decoders import SparseDepthWaveProgressiveDecoder   # noqa: E402

ks = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [9]
sh = 1
for a in sys.argv[2:]:
    if a.startswith("sh="):
        sh = int(a[3:])
lib = _lib.load()
lib.wmd_conv_tc_set_shared_taps(sh)
ch, n = synth.RESNET50_CH, int(os.environ.get("WMD_TRACE_BATCH", "32"))
dec = SparseDepthWaveProgressiveDecoder(np.array(ch))
synth.bench_kitti_params(dec)
dec = dec.cuda().eval()
feats = [f.cuda() for f in synth.bench_kitti_features(n, 320, 1024, ch)]
dec(feats, 0.05)
torch.cuda.synchronize()

state = {"i": 0}
real = ops.conv_rows
K, S = 256, 8


def report(info):
    torch.cuda.synchronize()
    buf = np.zeros(3 * S * K, dtype=np.int64)
    assert lib.wmd_debug_tc_trace(buf.ctypes.data_as(ctypes.c_void_p)) == 0
    t = buf.reshape(3, S, K)
    tb = np.zeros(64 * 12, dtype=np.int64)
    assert lib.wmd_debug_tc_tile_trace(tb.ctypes.data_as(ctypes.c_void_p)) == 0
    tt = tb.reshape(64, 12)
    nch = info["taps"] * (-(-info["c0"] // 32) + -(-info["c1"] // 32))
    rows = int(info["count"][0]) if info["count"] is not None else info["rows"]
    print("launch %d: taps %d cin %d+%d cout %d rows %d, %d chunks/tile, sh=%d" %
          (info["k"], info["taps"], info["c0"], info["c1"], info["cout"], rows, nch, sh))
    # CTA 0's clock span over the launch's event time = the SM clock this kernel actually ran at
    last = 0
    for i in range(1, 64):                       # entries past this launch's tiles are left over from earlier launches
        if tt[i, 0] >= tt[last, 7] and tt[i, 7] > tt[i, 0]:
            last = i
        else:
            break
    span = tt[last, 7] - tt[0, 0]
    print("   launch %.1f us; CTA 0 first..last trace point %d clk -> SM clock >= %.2f GHz (traced tiles only: first %d)" %
          (state.get("us", 0.0), span, span / max(state.get("us", 1.0), 1e-9) / 1e3, 64))
    labels = ["tables", "first raw A", "chunk loop", "last epoch wait", "drain", "store", "end barrier"]
    for i in range(3):
        if tt[i, 7] <= tt[i, 0]:
            break
        d = [tt[i, j + 1] - tt[i, j] for j in range(7)]
        print("   tile %d: " % i + "  ".join("%s %d" % (l, v) for l, v in zip(labels, d)) + "   total %d clk" % (tt[i, 7] - tt[i, 0]))
        print("           tables = init %d + tap tables (thread 0) %d + wait for the other threads %d + slot tables %d" %
              (tt[i, 8] - tt[i, 0], tt[i, 9] - tt[i, 8], tt[i, 10] - tt[i, 9], tt[i, 1] - tt[i, 10]))
    hi = min(K, nch) - 2
    sel = [c for c in range(3, hi) if c % 32 not in (0, 1, 31)]
    if not sel:
        return
    for role, nm, lab in ((0, "split warp 0", ["wait raw A", "wait MMA(c-2)", "LDS+split+tcgen05.st", "wait::st + arrive"]),
                          (1, "gather warp 8", ["wait stage free", "issue fill", "-"]),
                          (2, "issuer 0", ["wait weight image (B)", "wait split A", "issue MMAs + commit"])):
        x = t[role]
        period = np.mean([x[0, c + 1] - x[0, c] for c in sel])
        parts = ["%s %d" % (lab[i], np.mean([x[i + 1, c] - x[i, c] for c in sel])) for i in range(len(lab))]
        print("   %-14s period %6.0f clk/chunk | %s" % (nm, period, "  ".join(parts)))
        if role in (0, 2) and info["taps"] == 9:
            for m in range(3):
                selm = [c for c in sel if c % 3 == m]
                parts = ["%s %d" % (lab[i], np.mean([x[i + 1, c] - x[i, c] for c in selm])) for i in range(len(lab))]
                print("        dx=%d: period %6.0f | %s" % (m - 1, np.mean([x[0, c + 1] - x[0, c] for c in selm]), "  ".join(parts)))


def hooked(x0, c0, wpacked, bias, cout, n_, h, w, **kw):
    timed = wpacked.kind == "tc" and state["i"] in ks
    if timed:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    y = real(x0, c0, wpacked, bias, cout, n_, h, w, **kw)
    if timed:
        e1.record()
        torch.cuda.synchronize()
        state["us"] = e0.elapsed_time(e1) * 1e3
    if wpacked.kind == "tc":
        k = state["i"]
        state["i"] += 1
        if k in ks:
            report(dict(k=k, taps=kw.get("taps", 9), c0=c0, c1=kw.get("c1", 0), cout=cout, count=kw.get("count"), rows=n_ * h * w))
    return y


ops.conv_rows = hooked
import wavelet_monodepth_b200.kitti_decoders as kd   # noqa: E402
kd.ops.conv_rows = hooked
dec(feats, 0.05)
torch.cuda.synchronize()
