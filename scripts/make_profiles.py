#!/usr/bin/env python
"""Turn one scripts/gpu_check.sh visit (gpurun_out/*_<tag>.*) into the tracked summaries under profiles/:
   python scripts/make_profiles.py <tag> <round-prefix, e.g. r01>"""
import csv, gzip, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rnd = sys.argv[1], sys.argv[2]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# 1. the bench line
bench = json.load(open(os.path.join(G, "bench_%s.json" % tag)))
json.dump(bench, open(os.path.join(P, "%s_bench_n1.json" % rnd), "w"), indent=1)

# 2. the ncu launch list of the bench command: raw (gzip) + per-kernel aggregate
src = os.path.join(G, "launches_%s.csv" % tag)
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = {}
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000.0 if r[ui] == "ns" else (v * 1000.0 if r[ui] == "ms" else v)
    a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, "%s_ncu_launches_summary.txt" % rnd), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 python bench.py --steps 1 --warmup 3 --no-e2e --cpu-sample-pairs 0\n")
    f.write("# per-launch times are cold-cache and serialised: the SHARE of a kernel is what compares with bench.py's live measurement\n")
    f.write("# (bench.py roofline.share_of_kernel_time = %.3f for k_icp_forward + k_icp_reverse)\n" % (bench["roofline"]["share_of_kernel_time"] or 0))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write("%-64s launches %5d  total %10.1f us  avg %8.2f us  share %.3f\n" % (k[:64], a[0], a[1], a[1] / a[0], a[1] / tot))
    icp = sum(a[1] for k, a in agg.items() if "k_icp_" in k)
    f.write("k_icp_forward + k_icp_reverse share: %.3f of %.1f us over %d launches\n" % (icp / tot, tot, sum(a[0] for a in agg.values())))
with open(src, "rb") as fi, gzip.open(os.path.join(P, "%s_ncu_launches.csv.gz" % rnd), "wb") as fo:
    shutil.copyfileobj(fi, fo)

# 3. the full capture of the dominant kernels: raw-page metrics that the design argues from
rep = os.path.join(G, "prof_%s.ncu-rep" % tag)
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(out.splitlines()))
hh, uu = rr[0], rr[1]
traffic = 0.0
with open(os.path.join(P, "%s_icp_fused_ncu_full.txt" % rnd), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:k_icp -s 200 -c 2 python bench.py --steps 1 --warmup 3 --no-e2e --cpu-sample-pairs 0\n")
    f.write("# one forward + one reverse launch of all 24 pairs (24 views x 200k points), mid-align; caches flushed by ncu before each replay\n")
    for r in rr[2:]:
        f.write("== %s\n" % r[hh.index("Kernel Name")])
        for k in KEYS:
            if k in hh:
                f.write("   %-84s %s %s\n" % (k, r[hh.index(k)], uu[hh.index(k)]))
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v, u = float(r[hh.index(k)].replace(",", "")), uu[hh.index(k)]
            traffic += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    f.write("== DRAM traffic of the launch pair (one iteration of %d pairs): %.1f MB; algorithmic bytes: %.1f MB\n" % (bench["roofline"]["pairs_per_launch"], traffic / 1e6, bench["roofline"]["bytes_per_launch"] / 1e6))
print("profiles written; dram traffic per launch pair %.0f bytes" % traffic)
