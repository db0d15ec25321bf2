"""Sharding of a turntable ring over ranks and the gather of per-pair results (host side, backend agnostic).

The shardable unit is one ring pair p: view (p+1) % V aligned onto view p (the ring edges of the reference's
registrationLUM / computeError, mvr/src/registrator.cpp:482-487, 640-651).  Pairs are block-partitioned over the
ranks; every rank aligns its block with the C++ driver, the fixed-size per-pair records are all-gathered (one
exchange of ~100 B per pair: NCCL on GPUs, gloo in the CPU tests), and every rank closes the ring on the host.
No point data crosses ranks."""
import numpy as np

REC = 24   # floats per pair record: pose[16] (column-major), n_corr, mse, iterations, status, queries_lo, queries_hi, pad, pad


def pair_range(rank, world, n_pairs):
    return (rank * n_pairs) // world, ((rank + 1) * n_pairs) // world


def views_needed(p0, p1, n_views):
    return sorted({v % n_views for p in range(p0, p1) for v in (p, p + 1)})


def pack_reports(reports, p0, p1):
    """reports: list indexed by pair (dicts with pose 4x4, n_corr, mse, iterations, status, nn_queries)."""
    out = np.zeros((p1 - p0, REC), dtype=np.float32)
    for k, p in enumerate(range(p0, p1)):
        r = reports[p]
        out[k, :16] = np.ascontiguousarray(np.asarray(r["pose"], dtype=np.float32).T).reshape(16)
        out[k, 16] = r["n_corr"]
        out[k, 17] = r["mse"]
        out[k, 18] = r["iterations"]
        out[k, 19] = r["status"]
        q = int(r.get("nn_queries", 0))
        out[k, 20] = q & 0xFFFFFF          # float32 holds 24 bits exactly
        out[k, 21] = (q >> 24) & 0xFFFFFF
    return out


def unpack_records(rec):
    rec = np.asarray(rec, dtype=np.float32).reshape(-1, REC)
    out = []
    for r in rec:
        out.append(dict(pose=r[:16].reshape(4, 4).T.copy(), n_corr=int(r[16]), mse=float(r[17]), iterations=int(r[18]),
                        status=int(r[19]), nn_queries=int(r[20]) | (int(r[21]) << 24)))
    return out


def gather_records(mine, rank, world, n_pairs, dist=None, device=None):
    """All-gather the per-pair records: `mine` = (p1 - p0) x REC float32 of this rank -> n_pairs x REC on every rank.
    `dist` = torch.distributed (initialised) when world > 1."""
    if world == 1:
        return np.asarray(mine, dtype=np.float32).reshape(n_pairs, REC)
    import torch
    block = max(pair_range(r, world, n_pairs)[1] - pair_range(r, world, n_pairs)[0] for r in range(world))
    pad = torch.zeros(block * REC, dtype=torch.float32)
    flat = torch.from_numpy(np.ascontiguousarray(mine, dtype=np.float32).reshape(-1))
    pad[:flat.numel()] = flat
    pad = pad.to(device) if device is not None else pad
    out = torch.empty(world * block * REC, dtype=torch.float32, device=pad.device)
    if hasattr(dist, "all_gather_into_tensor") and pad.is_cuda:
        dist.all_gather_into_tensor(out, pad)          # one exchange, one device-to-host read
    else:                                              # gloo (CPU tests)
        parts = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        out = torch.cat(parts)
    host = out.cpu().numpy().reshape(world, block, REC)
    rows = []
    for r in range(world):
        a, b = pair_range(r, world, n_pairs)
        rows.append(host[r, :b - a])
    return np.concatenate(rows, axis=0)


def close_ring(allrec, centre, radius, relax=True, iterations=16):
    """Loop closure over the gathered records (every rank computes the same poses)."""
    from . import ring_close
    recs = unpack_records(allrec)
    rel = [r["pose"] for r in recs]
    w = [float(r["n_corr"]) if r["status"] == 0 else 0.0 for r in recs]
    return ring_close(rel, w, relax=relax, iterations=iterations, centre=centre, rot_scale=radius)
