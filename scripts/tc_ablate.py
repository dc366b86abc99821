"""Timing ablations of conv_rows_tc (WMD_TC_EXP = k builds; their RESULTS ARE WRONG, only the clock matters).

    python scripts/tc_ablate.py build 1 2 3 5 6    # here: scripts/bench_cu/_bin/libwmd_exp<k>.so
    WMD_LIB_PATH=scripts/bench_cu/_bin/libwmd_exp1.so WMD_CONV_PRECISION=f16x3 python scripts/conv_layers_env.py   # GPU box

 f16 operand form (WMD_CONV_PRECISION=f16x3):
   1 no conversion arithmetic in the split     2 = 1 + no shared-memory row reads     3 = 2 + no tcgen05.wait::st
 epilogue (either form):
   5 output stores go to a 2 MB window that stays in L2     6 no output stores
 Only the dense level-4 layers keep their shape under an ablation (later masks depend on the wrong outputs).
 Measured (B200, R50 1024x320 bs32): upconv(4,1) f16x3 827 us -> 713 (1) -> 641 (2) = (3); tf32x3 920.  1x1 256->576: 134 us ->
 134 (5) -> 104 (6): the store instructions cost 22 % of that layer wherever they land, DRAM is not what they wait for.
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
