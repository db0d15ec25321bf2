"""Sharding of a turntable ring over ranks and the gather of per-pair results (host side, backend agnostic).

The shardable unit is one ring pair p: view (p+1) % V aligned onto view p (the ring edges of the reference's
registrationLUM / computeError, mvr/src/registrator.cpp:482-487, 640-651).  Pairs are block-partitioned over the
ranks; every rank aligns its block with the C++ driver, the fixed-size per-pair records are all-gathered (one
exchange of ~100 B per pair: NCCL on GPUs, gloo in the CPU tests), and every rank closes the ring on the host.
No point data crosses ranks."""
import numpy as np

# One pair record = the C struct mvr_pair_record of include/mvr_b200.h (96 bytes, no padding): what a rank contributes to
# the all-gather.  Carried as raw bytes, so n_corr (int32), mse (double) and the query count (uint64) arrive exactly
# as the single-GPU path reports them.
RECORD = np.dtype([("pose", np.float32, 16), ("n_corr", np.int32), ("iterations", np.int32), ("status", np.int32), ("pad", np.int32),
                   ("mse", np.float64), ("nn_queries", np.uint64)])
REC = RECORD.itemsize   # bytes per pair record
assert REC == 96


def pair_range(rank, world, n_pairs):
    return (rank * n_pairs) // world, ((rank + 1) * n_pairs) // world


def views_needed(p0, p1, n_views):
    return sorted({v % n_views for p in range(p0, p1) for v in (p, p + 1)})


def pack_reports(reports, p0, p1):
    """reports: list indexed by pair (dicts with pose 4x4, n_corr, mse, iterations, status, nn_queries)."""
    out = np.zeros(p1 - p0, dtype=RECORD)
    for k, p in enumerate(range(p0, p1)):
        r = reports[p]
        out[k]["pose"] = np.ascontiguousarray(np.asarray(r["pose"], dtype=np.float32).T).reshape(16)   # column-major
        out[k]["n_corr"] = r["n_corr"]
        out[k]["iterations"] = r["iterations"]
        out[k]["status"] = r["status"]
        out[k]["mse"] = r["mse"]
        out[k]["nn_queries"] = int(r.get("nn_queries", 0))
    return out


def records_from_reports(reports, p0, p1):
    """The same from the raw report array of Registrator.register_turntable(raw=True): no per-pair Python objects."""
    r = reports[p0:p1]
    out = np.zeros(p1 - p0, dtype=RECORD)
    out["pose"] = r["pose"]
    out["n_corr"] = r["n_correspondences"]
    out["iterations"] = r["iterations"]
    out["status"] = r["status"]
    out["mse"] = r["mse"]
    out["nn_queries"] = r["nn_queries"]
    return out


def unpack_records(rec):
    rec = np.asarray(rec).view(RECORD).reshape(-1)
    return [dict(pose=r["pose"].reshape(4, 4).T.copy(), n_corr=int(r["n_corr"]), mse=float(r["mse"]), iterations=int(r["iterations"]),
                 status=int(r["status"]), nn_queries=int(r["nn_queries"])) for r in rec]


def pose_checksum(allrec):
    """Hash of the bit patterns of every gathered pair pose (plus n_corr and mse): equal checksums = bit-identical results,
    whatever the number of ranks."""
    import hashlib
    rec = np.asarray(allrec).view(RECORD).reshape(-1)
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(rec["pose"]).tobytes())
    h.update(np.ascontiguousarray(rec["n_corr"]).tobytes())
    h.update(np.ascontiguousarray(rec["mse"]).tobytes())
    return h.hexdigest()[:16]


_GATHER_CACHE = {}


def gather_records(mine, rank, world, n_pairs, dist=None, device=None):
    """All-gather the per-pair records: `mine` = the (p1 - p0) RECORDs of this rank -> n_pairs RECORDs on every rank.
    `dist` = torch.distributed (initialised) when world > 1."""
    mine = np.ascontiguousarray(np.asarray(mine).view(RECORD).reshape(-1))
    if world == 1:
        assert len(mine) == n_pairs
        return mine
    import torch
    block = max(pair_range(r, world, n_pairs)[1] - pair_range(r, world, n_pairs)[0] for r in range(world))
    if device is not None and str(device).startswith("cuda") and hasattr(dist, "all_gather_into_tensor"):
        # NCCL: staging buffers made once (pinned on the host), one copy in, one exchange, one copy out
        key = (world, n_pairs, str(device))
        c = _GATHER_CACHE.get(key)
        if c is None:
            c = dict(h_in=torch.zeros(block * REC, dtype=torch.uint8).pin_memory(), d_in=torch.zeros(block * REC, dtype=torch.uint8, device=device),
                     d_out=torch.empty(world * block * REC, dtype=torch.uint8, device=device),
                     h_out=torch.empty(world * block * REC, dtype=torch.uint8).pin_memory())
            _GATHER_CACHE[key] = c
        flat = mine.view(np.uint8).reshape(-1)
        c["h_in"].numpy()[:flat.size] = flat
        c["d_in"].copy_(c["h_in"], non_blocking=True)
        dist.all_gather_into_tensor(c["d_out"], c["d_in"])
        c["h_out"].copy_(c["d_out"], non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
        host = c["h_out"].numpy().reshape(world, block * REC)
        rows = []
        for r in range(world):
            a, b = pair_range(r, world, n_pairs)
            rows.append(host[r, :(b - a) * REC].copy().view(RECORD))
        return np.concatenate(rows, axis=0)
    pad = torch.zeros(block * REC, dtype=torch.uint8)
    flat = torch.from_numpy(mine.view(np.uint8).reshape(-1).copy())
    pad[:flat.numel()] = flat
    pad = pad.to(device) if device is not None else pad
    out = torch.empty(world * block * REC, dtype=torch.uint8, device=pad.device)
    if hasattr(dist, "all_gather_into_tensor") and pad.is_cuda:
        dist.all_gather_into_tensor(out, pad)          # one exchange, one device-to-host read
    else:                                              # gloo (CPU tests)
        parts = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        out = torch.cat(parts)
    host = out.cpu().numpy().reshape(world, block * REC)
    rows = []
    for r in range(world):
        a, b = pair_range(r, world, n_pairs)
        rows.append(host[r, :(b - a) * REC].copy().view(RECORD))
    return np.concatenate(rows, axis=0)


def close_ring(allrec, centre, radius, relax=True, iterations=16):
    """Loop closure over the gathered records (every rank computes the same poses)."""
    from . import ring_close_flat
    rec = np.asarray(allrec).view(RECORD).reshape(-1)
    w = np.where(rec["status"] == 0, rec["n_corr"].astype(np.float64), 0.0)
    return ring_close_flat(np.ascontiguousarray(rec["pose"]), w, relax=relax, iterations=iterations, centre=centre, rot_scale=radius)
