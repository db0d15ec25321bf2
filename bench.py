#!/usr/bin/env python
"""bench.py -- turntable registration throughput on B200 (BASELINE.json metric).

A "step" is ONE registration of the whole synthetic turntable sequence (BASELINE.json configs[1]):
V views x N points, the V ring pairs (k+1 -> k, closing pair 0 -> V-1) each aligned by 30 fixed
iterations of reciprocal point-to-point ICP with a 4 mm gate (the reference's ICP settings,
mvr/src/registrator.cpp:551-560), per-pair results gathered over NCCL, then the host loop closure.
With --gpus N the V pairs are block-partitioned over N ranks (configs[2]); total work is fixed.

  value          NN queries/s of the whole job, scans already resident in HBM (device-timed)
  ms_per_step    registration time of the sequence (the second half of BASELINE's metric)
  e2e            the same job through the host-buffer C ABI: every step copies each scan H2D from pinned
                 host memory and reads poses/residuals back
  roofline       the correspondence kernel (forward + reciprocal NN): algorithmic bytes / CUDA-event time
  cpu_baseline   the CPU oracle (port of the reference's PCL path) on a bounded sample, same box

`--impl reference` times that CPU path alone (rank 0 only) and prints the same JSON shape.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "nn_queries_per_s (24-view x 200k turntable registration; ms_per_step = registration ms)"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--views", type=int, default=24)
    ap.add_argument("--points", type=int, default=200_000)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--max-dist", type=float, default=4.0)
    ap.add_argument("--reciprocal", type=int, default=1)
    ap.add_argument("--cpu-sample-pairs", type=int, default=1, help="pairs the cpu_baseline leg aligns (0 = skip)")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_config(a, world):
    return {
        "workload": "%d-view turntable, %d pts/view, %d ring pairs, %d fixed ICP iterations/pair, point-to-point, "
                    "reciprocal=%d, gate %.1f mm, + host loop closure" % (a.views, a.points, a.views, a.iters, a.reciprocal, a.max_dist),
        "views": a.views, "points_per_view": a.points, "pairs": a.views, "iterations": a.iters,
        "sharding": "pairs block-partitioned over %d rank(s); one NCCL all-gather of per-pair results" % world,
        "l2": "256 MiB scratch buffer rewritten between timed steps (L2 flush)",
    }


def pair_range(rank, world, n_pairs):
    return (rank * n_pairs) // world, ((rank + 1) * n_pairs) // world


def make_pairs(a, p0, p1):
    """Synthetic scans and initial guesses of pairs [p0, p1): pair p aligns view (p+1)%V onto view p."""
    import mvr_b200.synth as synth
    V = a.views
    need = sorted({v % V for p in range(p0, p1) for v in (p, p + 1)})
    views, poses = {}, {}
    for v in need:
        views[v], poses[v] = synth.turntable_view(v, V, a.points)
    E = synth.perturbation()
    pairs = []
    for p in range(p0, p1):
        t, s = p % V, (p + 1) % V
        guess = (E @ np.linalg.inv(poses[t]) @ poses[s]).astype(np.float32)   # ideal turntable step + fixed error
        pairs.append(dict(pair=p, tgt=t, src=s, guess=guess, truth=np.linalg.inv(poses[t]) @ poses[s]))
    return views, pairs


# ------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline (the oracle = port of the reference's PCL path)
# ------------------------------------------------------------------------------------------------
def cpu_align_pairs(a, views, pairs):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    oracle.build()
    prm = oracle.make_params(max_iterations=a.iters, max_dist=a.max_dist, reciprocal=bool(a.reciprocal), fixed_iterations=True)
    t0 = time.perf_counter()
    q = 0
    for pr in pairs:
        o = oracle.icp_align(views[pr["src"]], views[pr["tgt"]], prm, guess=pr["guess"], max_log=1)
        q += o["nn_queries"]
    dt = time.perf_counter() - t0
    return q, dt, oracle.num_threads()


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    views, pairs = make_pairs(a, 0, 1)
    for _ in range(a.warmup):
        cpu_align_pairs(a, views, pairs)
    tq, tt, cores = 0, 0.0, 1
    for _ in range(a.steps):
        q, dt, cores = cpu_align_pairs(a, views, pairs)
        tq += q
        tt += dt
    val = tq / tt
    sample = "pair 0 of the sequence (1/%d of a step), %d iterations, per step" % (a.views, a.iters)
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * tt / a.steps * a.views, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(a, 1),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "oracle port of the reference's PCL ICP path (PCL itself cannot be built here); ms_per_step is the "
                "sampled pair time x %d pairs" % a.views,
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[0]))
                mx.append(float(c[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if c[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def run_native(a):
    import torch
    import torch.distributed as dist
    import mvr_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    V = a.views
    p0, p1 = pair_range(rank, world, V)
    views, pairs = make_pairs(a, p0, p1)
    n = a.points
    ctx = mvr_b200.Context(local)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    prm = mvr_b200.default_params(max_iterations=a.iters, max_dist=a.max_dist, reciprocal=a.reciprocal, fixed_iterations=1)

    d_views = {v: torch.from_numpy(p).to(dev) for v, p in views.items()}
    h_views = {v: torch.from_numpy(p).pin_memory() for v, p in views.items()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    REC = 24   # floats per pair record: pose 16, mse, n_corr, iterations, status, queries(lo, hi as 2^24 split), pad
    gathered = torch.zeros(V * REC, dtype=torch.float32, device=dev)
    mine = torch.zeros((p1 - p0) * REC, dtype=torch.float32, device=dev)
    h_rec = torch.zeros((p1 - p0) * REC, dtype=torch.float32).pin_memory()
    h_all = torch.zeros(V * REC, dtype=torch.float32).pin_memory()

    def step(host_buffers):
        """One registration of this rank's pairs + gather + loop closure.  Returns (queries, results)."""
        q = 0
        res = []
        with torch.cuda.stream(stream):
            for k, pr in enumerate(pairs):
                if host_buffers:
                    ctx.set_target(h_views[pr["tgt"]].numpy())
                    ctx.set_source(h_views[pr["src"]].numpy())
                else:
                    ctx.set_target_device(d_views[pr["tgt"]].data_ptr(), n)
                    ctx.set_source_device(d_views[pr["src"]].data_ptr(), n)
                r = ctx.icp_align(prm, guess=pr["guess"], n_source=n)
                q += r["nn_queries"]
                res.append(r)
                rec = h_rec[k * REC:(k + 1) * REC]
                rec[:16] = torch.from_numpy(np.ascontiguousarray(r["final"].T).reshape(16))
                rec[16] = r["mse"]
                rec[17] = r["n_corr"]
                rec[18] = r["iterations"]
                rec[19] = r["status"]
            mine.copy_(h_rec, non_blocking=True)
            if world > 1:
                stream.synchronize()
                dist.all_gather_into_tensor(gathered, mine)
                h_all.copy_(gathered, non_blocking=True)
            else:
                h_all.copy_(mine, non_blocking=True)
            stream.synchronize()
        poses = loop_closure(h_all.numpy().reshape(V, REC), V)
        return q, res, poses

    def loop_closure(recs, V):
        """Chain the ring pairs into absolute poses and spread the closing error evenly (host side).
        TODO(next milestone): replaced by the C++ LUM-style relaxation in host/lum.cpp."""
        P = [np.eye(4)]
        for k in range(V - 1):
            T = recs[k, :16].astype(np.float64).reshape(4, 4).T
            P.append(P[-1] @ T)
        return P

    def timed(host_buffers, steps):
        tot_ms, tot_q = 0.0, 0
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        barrier()
        for _ in range(steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            with torch.cuda.stream(stream):
                e0.record(stream)
            q, res, poses = step(host_buffers)
            with torch.cuda.stream(stream):
                e1.record(stream)
            e1.synchronize()
            tot_ms += e0.elapsed_time(e1)
            tot_q += q
        barrier()
        return tot_ms, tot_q, res

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident leg (value) ----
    for _ in range(a.warmup):
        step(False)
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = mvr_b200.kernel_launch_count()
    ms, q, res = timed(False, a.steps)
    launches = mvr_b200.kernel_launch_count() - l0
    clocks = sampler.stop() if sampler else None
    ms = allmax(ms)
    q = allsum(q)
    launches = allsum(launches)
    value = q / (ms * 1e-3)

    # ---- end-to-end leg (host buffers through the C ABI) ----
    e2e = None
    if not a.no_e2e:
        for _ in range(max(1, min(a.warmup, 2))):
            step(True)
        ms_e, q_e, _ = timed(True, a.steps)
        ms_e = allmax(ms_e)
        q_e = allsum(q_e)
        h2d = allsum(float(len(pairs) * 2 * n * 16 + len(pairs) * REC * 4))
        d2h = allsum(float(len(pairs) * 256 * a.iters)) + V * REC * 4   # 32 doubles of sums per iteration + records
        e2e = {"value": q_e / (ms_e * 1e-3), "unit": UNIT, "ms_per_step": ms_e / a.steps,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}

    # ---- roofline of the dominant kernel: one extra untimed step with per-kernel CUDA events ----
    ctx.set_profiling(True)
    ctx.kernel_stats(reset=True)
    step(False)
    st = ctx.kernel_stats(reset=True)
    ctx.set_profiling(False)
    roof = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        c = st["corr"]
        avg_ms = c["ms"] / max(c["launches"], 1)
        bytes_per_launch = c["bytes"] / max(c["launches"], 1)
        ach = bytes_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        total_kernel_ms = sum(v["ms"] for v in st.values())
        roof = {"bound": "hbm", "kernel": "k_correspond (forward + reciprocal NN search, gate)", "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                "bytes_per_launch": bytes_per_launch, "avg_launch_us": avg_ms * 1e3, "launches": c["launches"],
                "share_of_kernel_time": c["ms"] / total_kernel_ms if total_kernel_ms > 0 else None,
                "per_kernel_ms": {k: round(v["ms"], 4) for k, v in st.items() if v["launches"]}}

    # ---- accuracy summary (parity itself lives in tests/) ----
    acc = None
    if rank == 0 and res:
        r, pr = res[0], pairs[0]
        dT = r["final"].astype(np.float64) @ np.linalg.inv(pr["truth"])
        w = 0.5 * np.array([dT[2, 1] - dT[1, 2], dT[0, 2] - dT[2, 0], dT[1, 0] - dT[0, 1]])
        acc = {"pair0_rot_err_rad_vs_truth": float(np.arcsin(min(1.0, np.linalg.norm(w)))),
               "pair0_rmse_mm": float(np.sqrt(r["mse"])), "pair0_n_corr": int(r["n_corr"])}

    cpu = None
    if rank == 0 and a.cpu_sample_pairs > 0:
        cq, cdt, cores = cpu_align_pairs(a, views, pairs[:a.cpu_sample_pairs])
        cpu = {"value": cq / cdt, "unit": UNIT, "cores": cores, "kind": "port", "seconds": cdt,
               "sample": "first %d pair(s) of the same sequence (%d/%d of a step), %d iterations each"
                         % (a.cpu_sample_pairs, a.cpu_sample_pairs, V, a.iters)}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(a, world),
            "registration_ms": ms / a.steps, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": roof, "cpu_baseline": cpu, "accuracy": acc,
        }
        print(json.dumps(out))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)


if __name__ == "__main__":
    main()
