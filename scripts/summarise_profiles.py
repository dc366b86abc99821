"""Turn the ncu outputs of scripts/gpu_profile.sh (gpurun_out/launches.csv, gpurun_out/prof_conv.ncu-rep) into the small
tracked summaries under profiles/:  python scripts/summarise_profiles.py <tag>   (needs `ncu` on PATH, no GPU)."""
import csv
import io
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(REPO, "profiles")

# ---- 1. launch list: second decoder step, per kernel totals
rows = []
with open(os.path.join(REPO, "gpurun_out", "launches.csv")) as f:
    text = "".join(l for l in f if l.startswith('"'))
for r in csv.DictReader(io.StringIO(text)):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"])))
# profile_step.py ran 2 decoder steps; keep the second (weights packed, allocator warm).  A step starts with the layout
# move of the coarsest feature map and has five of them: the second step starts at the sixth nchw_to_rows launch.
# (the number of whole-map moves per step depends on the layout options: take the second half of them)
moves = [i for i, (name, _) in enumerate(rows) if "nchw_to_rows" in name]
step = rows[moves[len(moves) // 2]:] if len(moves) >= 2 and len(moves) % 2 == 0 else rows[len(rows) // 2:]
agg = {}
for name, ns in step:
    short = name.split("(")[0].replace("void ", "").replace("wmd::", "")
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += ns
total = sum(v[1] for v in agg.values())
with open(os.path.join(out_dir, "%s_launches_summary.csv" % tag), "w") as f:
    f.write("# %s: ncu --metrics gpu__time_duration.sum --clock-control none python scripts/profile_step.py 2 ; second decoder "
            "step (R50 1024x320 bs32), %d launches, %.0f us; per-launch times are cold-cache and serialised\n" % (tag, len(step), total / 1e3))
    f.write("kernel,launches,total_ns,share\n")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%s,%d,%d,%.4f\n" % (k, n, ns, ns / total))

# ---- 2. full-set capture of the conv kernel: a few columns per launch
raw = subprocess.run(["ncu", "-i", os.path.join(REPO, "gpurun_out", "prof_conv.ncu-rep"), "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rd = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rd[0], rd[1], rd[2:]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
idx = [hdr.index(w) for w in want if w in hdr]
with open(os.path.join(out_dir, "%s_conv_rows_tc_ncu_full.csv" % tag), "w") as f:
    f.write("# %s: ncu --set full --clock-control none --import-source on -k regex:conv_rows -s 12 -c 12 python "
            "scripts/profile_step.py 2 (second decoder step, R50 1024x320 bs32)\n" % tag)
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in data:
        w.writerow([r[i] for i in idx])
ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


tc = [to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]) for r in data if "conv_rows_tc" in r[it]]
traffic_path = os.path.join(out_dir, "ncu_traffic.json")
tj = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
if tc:
    tj["conv_rows_tc"] = int(sum(tc) / len(tc))
    tj["source"] = ("profiles/%s_conv_rows_tc_ncu_full.csv (mean dram read+write over the %d conv_rows_tc launches of one decoder "
                    "step) and profiles/r01_conv_rows_ncu_full.csv (SIMT engine)" % (tag, len(tc)))
    json.dump(tj, open(traffic_path, "w"))
print("launches:", len(step), "conv launches in capture:", len(data), "mean conv traffic:", tj.get("conv_rows_tc"))

# ---- 3. second capture: the HBM-bound kernels (fused level tail = IDWT chain, layout moves)
other = os.path.join(REPO, "gpurun_out", "prof_other.ncu-rep")
if os.path.exists(other):
    raw = subprocess.run(["ncu", "-i", other, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rd[0], rd[1], rd[2:]
    idx = [hdr.index(w) for w in want if w in hdr]
    with open(os.path.join(out_dir, "%s_idwt_layout_ncu_full.csv" % tag), "w") as f:
        f.write("# %s: ncu --set full --clock-control none --import-source on -k regex:'head_idwt|nchw_to_rows|gather_rows_list' -s 9 -c 9 python "
                "scripts/profile_step.py 2 (second decoder step, R50 1024x320 bs32): head_idwt = fused level tail (head gather-sum -> "
                "yh -> IDWT -> disp -> next threshold), nchw_to_rows = whole-map layout moves (dense level), gather_rows_list = "
                "compact skip rows of the sparse levels\n" % tag)
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in data:
            w.writerow([r[i] for i in idx])
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    for key in ("head_idwt", "nchw_to_rows", "gather_rows_list"):
        v = [to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]) for r in data if key in r[it]]
        if v:
            tj[key] = int(sum(v) / len(v))
    tj["source_other"] = "profiles/%s_idwt_layout_ncu_full.csv (mean dram read+write per launch)" % tag
    json.dump(tj, open(traffic_path, "w"))
    print("second capture:", len(data), "launches;", {k: tj.get(k) for k in ("head_idwt", "nchw_to_rows", "gather_rows_list")})
