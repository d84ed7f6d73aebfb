#!/usr/bin/env python
"""Summarise gpurun_out/<tag>/{launches.csv,prof_*.ncu-rep,bench.json} into profiles/<tag>_*.txt (read here, no GPU)."""
import collections, csv, json, os, subprocess, sys

tag = sys.argv[1]
src = os.path.join("gpurun_out", tag)
os.makedirs("profiles", exist_ok=True)
out = []
lf = os.path.join(src, "launches.csv")
if os.path.isfile(lf):
    rows = list(csv.reader(open(lf)))
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                v = float(d["Metric Value"].replace(",", ""))
                v = v / 1e6 if d["Metric Unit"] == "ns" else v / 1e3 if d["Metric Unit"] == "us" else v
                agg[d["Kernel Name"][:70]][0] += 1
                agg[d["Kernel Name"][:70]][1] += v
    tot = sum(v[1] for v in agg.values())
    out.append("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)")
    fam = collections.defaultdict(float)
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append("%-72s n=%4d total=%9.3f ms avg=%8.4f ms share=%5.1f%%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
        fam[k.split("<")[0].replace("void ", "").split("(")[0]] += v[1]
    out.append("# by kernel family (the list covers the whole bench.py process: k_any_nonzero = grid-class check at create time,")
    out.append("# k_repitch = set/get_wavefield of the e2e leg, k_pack_box = snapshot frames of the e2e leg; a device-resident step is")
    out.append("# 4 x (k_top + k_main_tma + k_src_inject) + k_pg + k_record)")
    for k, v in sorted(fam.items(), key=lambda x: -x[1]):
        out.append("%-40s total=%9.3f ms share=%5.1f%%" % (k, v, 100 * v / tot))
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
def raw_rows(path_rep=None, path_csv=None):
    if path_csv:
        return list(csv.reader(open(path_csv)))
    p = subprocess.run(["ncu", "-i", path_rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    return list(csv.reader(p.stdout.splitlines()))


traffic = []
for f in sorted(os.listdir(src)):
    # the reports are exported to CSV on the GPU box (scripts/gpu_ncu.sh): a .ncu-rep with sources is too large to travel
    if f.endswith(".ncu-rep") or f.endswith("_raw.csv"):
        rows = raw_rows(path_rep=os.path.join(src, f)) if f.endswith(".ncu-rep") else raw_rows(path_csv=os.path.join(src, f))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        out.append("\n# ncu --set full: %s" % f)
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            out.append(d["Kernel Name"][:80])
            for w in want:
                if w in d:
                    out.append("    %-90s %s %s" % (w, d[w], units[hdr.index(w)]))
            if "k_main_tma" in d["Kernel Name"]:
                def gb(k):
                    v = float(d[k]); u = units[hdr.index(k)]
                    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
                traffic.append((d["Kernel Name"], gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")))
if traffic:
    out.append("\n# DRAM traffic (read + write) per captured launch of the dominant kernel")
    for k, b in traffic:
        out.append("%-70s %.3f GB" % (k[:70], b / 1e9))
bj = os.path.join(src, "bench.json")
if os.path.isfile(bj):
    out.append("\n# bench.py line of the same session")
    out.append(open(bj).read().strip())
open(os.path.join("profiles", tag + "_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[-60:]))
