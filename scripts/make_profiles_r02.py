#!/usr/bin/env python
"""Tracked summaries under profiles/ from one measurement visit (round 2 layout: two ncu reports, one per kernel):
   python scripts/make_profiles_r02.py <tag of gpurun_out files, e.g. r02v> <round prefix, e.g. r02>"""
import csv, gzip, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rnd = sys.argv[1], sys.argv[2]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
bench = json.loads(open(os.path.join(G, "%s_bench_n1.json" % tag)).read().strip().splitlines()[-1])
json.dump(bench, open(os.path.join(P, "%s_bench_n1.json" % rnd), "w"), indent=1)
ref = os.path.join(G, "%s_ref.json" % tag)
if os.path.exists(ref):
    json.dump(json.loads(open(ref).read().strip().splitlines()[-1]), open(os.path.join(P, "%s_bench_reference_arm.json" % rnd), "w"), indent=1)

src = os.path.join(G, "%s_launches.csv" % tag)
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
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 python bench.py --steps 2 --warmup 1 --no-e2e --cpu-sample-pairs 0\n")
    f.write("# per-launch times are cold-cache and serialised: the SHARE of a kernel is what compares with bench.py's live measurement\n")
    f.write("# (bench.py roofline.share_of_step = %.3f: the CUDA-event span of the iterations / the step)\n" % (bench["roofline"].get("share_of_step") or 0))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write("%-64s launches %5d  total %10.1f us  avg %8.2f us  share %.3f\n" % (k[:64], a[0], a[1], a[1] / a[0], a[1] / tot))
    icp = sum(a[1] for k, a in agg.items() if "k_icp_" in k)
    f.write("k_icp_forward + k_icp_reverse share: %.3f of %.1f us over %d launches\n" % (icp / tot, tot, sum(a[0] for a in agg.values())))
with open(src, "rb") as fi, gzip.open(os.path.join(P, "%s_ncu_launches.csv.gz" % rnd), "wb") as fo:
    shutil.copyfileobj(fi, fo)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__shared_mem_per_block_static", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
traffic = 0.0
pairs = bench["config"]["pairs"]
n = bench["config"]["points_per_view"]
with open(os.path.join(P, "%s_icp_fused_ncu_full.txt" % rnd), "w") as f:
    f.write("# MVR_ITERS=30 ncu --set full --clock-control none --import-source on -k regex:k_icp_forward|k_icp_reverse -s 20 -c 1 python scripts/gpu_iter_profile.py\n")
    f.write("# (two captures, one per kernel; MVR batch group forced to 24) one forward + one reverse launch of all 24 pairs (24 views x 200k points), iteration 21 of 30; caches flushed by ncu before each replay\n")
    for which in ("fwd", "rev"):
        rep = os.path.join(G, "%s_%s.ncu-rep" % (tag, which))
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(out.splitlines()))
        hh, uu = rr[0], rr[1]
        r = rr[2]
        f.write("== %s\n" % r[hh.index("Kernel Name")])
        vals = {}
        for k in KEYS:
            if k in hh:
                f.write("   %-84s %s %s\n" % (k, r[hh.index(k)], uu[hh.index(k)]))
                vals[k] = r[hh.index(k)]
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v, u = float(r[hh.index(k)].replace(",", "")), uu[hh.index(k)]
            traffic += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        ti = float(vals["smsp__inst_executed.sum"].replace(",", "")) * float(vals["smsp__thread_inst_executed_per_inst_executed.ratio"].replace(",", ""))
        q = pairs * n if which == "fwd" else None
        if q:
            f.write("   thread-instructions per forward query (%d queries)                                  %.0f\n" % (q, ti / q))
    f.write("== DRAM traffic of the launch pair (one iteration of %d pairs): %.1f MB; algorithmic bytes (SURVEY.md 8d): %.1f MB; without the re-index bytes: %.1f MB\n"
            % (pairs, traffic / 1e6, bench["roofline"]["bytes_per_launch"] / 1e6, pairs * 56.0 * n / 1e6))
print("profiles written; dram traffic per launch pair %.0f bytes" % traffic)
