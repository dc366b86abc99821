"""Timing ablations of conv_rows_tc (WMD_TC_EXP = k builds; their RESULTS ARE WRONG, only the clock matters).

    python scripts/tc_ablate.py build 1 2 3 4      # here: scripts/bench_cu/_bin/libwmd_exp<k>.so
    WMD_LIB_PATH=scripts/bench_cu/_bin/libwmd_exp1.so WMD_CONV_PRECISION=f16x3 python scripts/conv_layers_env.py   # GPU box

 1 no conversion arithmetic in the f16 split     2 = 1 + no shared-memory row reads     3 = 2 + no tcgen05.wait::st
 4 issuers do not wait for the split (free-running MMAs: the tensor / weight-stream floor)
"""
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from wavelet_monodepth_b200 import build as wbuild   # noqa: E402

if sys.argv[1] == "flags":        # python scripts/tc_ablate.py flags <name> "<nvcc -D flags>"  ->  _bin/libwmd_<name>.so
    out = os.path.join(REPO, "scripts", "bench_cu", "_bin")
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, "libwmd_%s.so" % sys.argv[2])
    cmd = [wbuild.nvcc_path()] + sys.argv[3].split() + wbuild.NVCC_FLAGS + ["-I", os.path.join(REPO, "include"), "-I", wbuild.CSRC,
                                                                          "-o", lib] + wbuild.sources()
    subprocess.run(cmd, check=True)
    print(lib)
    sys.exit(0)

if sys.argv[1] == "build":
    out = os.path.join(REPO, "scripts", "bench_cu", "_bin")
    os.makedirs(out, exist_ok=True)
    procs = []
    for k in sys.argv[2:]:
        lib = os.path.join(out, "libwmd_exp%s.so" % k)
        cmd = [wbuild.nvcc_path(), "-DWMD_TC_EXP=%s" % k] + wbuild.NVCC_FLAGS + ["-I", os.path.join(REPO, "include"), "-I", wbuild.CSRC,
                                                                              "-o", lib] + wbuild.sources()
        procs.append((lib, subprocess.Popen(cmd)))
    for lib, pr in procs:
        assert pr.wait() == 0, lib
        print(lib)
