"""Multi-GPU plumbing: instances are independent, so a batch is split contiguously across ranks with no
data-path collective; one all-gather collects converged trajectories, duals, stats and status at the end."""
from __future__ import annotations

import numpy as np


def shard_bounds(total: int, world: int, rank: int):
    """Contiguous split of `total` instances: ranks < total % world get one extra."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _CudaArray:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr, "version": 2}


def device_tensor(ptr, shape, dtype="f8", device=0):
    """Zero-copy torch view of a device buffer owned by an agb_handle (agb_get_device_view)."""
    import torch
    return torch.as_tensor(_CudaArray(ptr, shape, "<" + dtype), device=f"cuda:{device}")


def result_views(batch):
    """torch views (no copy) of the resident results of a GameBatch."""
    v, B = batch.device_view(), batch.batch
    dev = batch.device
    out = {
        "Z": device_tensor(v.Z_dev, (B, batch.N * (batch.n + batch.m)), "f8", dev),
        "L": device_tensor(v.L_dev, (B, batch.p * (batch.N - 1) * batch.n), "f8", dev),
        "stats": device_tensor(v.stats_dev, (B, 10), "f8", dev),
        "status": device_tensor(v.status_dev, (B,), "i4", dev),
    }
    return out


def results_slab(batch):
    """Zero-copy FP64 view of the handle's whole result allocation [Z | L | stats | status(int32, padded)]."""
    v = batch.device_view()
    return device_tensor(v.results_dev, (v.results_bytes // 8,), "f8", batch.device)


def all_gather_slabs(local_flat, out=None, group=None):
    """The single collective of the path: every rank's result slab, rank-major, straight from the solver's buffers."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world * local_flat.numel(),), dtype=local_flat.dtype, device=local_flat.device)
    dist.all_gather_into_tensor(out, local_flat, group=group)
    return out


def unpack_slab(flat, B, N, n, m, p):
    """One rank's slab -> dict of numpy views."""
    a = flat.cpu().numpy() if hasattr(flat, "cpu") else np.asarray(flat)
    zs, ls, ss = B * N * (n + m), B * p * (N - 1) * n, B * 10
    return {"Z": a[:zs].reshape(B, N, n + m), "L": a[zs:zs + ls].reshape(B, p, N - 1, n),
            "stats": a[zs + ls:zs + ls + ss].reshape(B, 10), "status": a[zs + ls + ss:].view(np.int32)[:B]}


def pack_results(views, out=None):
    """[B, zs+ls+10+1] FP64 slab (status widened to FP64) — one buffer, so one collective."""
    import torch
    parts = [views["Z"], views["L"], views["stats"], views["status"].to(torch.float64)[:, None]]
    if out is None:
        return torch.cat(parts, dim=1)
    torch.cat(parts, dim=1, out=out)
    return out


def all_gather_results(local_slab, group=None):
    """The single collective of the path: all-gather of the per-rank result slabs (equal shard sizes)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty((world * local_slab.shape[0], local_slab.shape[1]), dtype=local_slab.dtype, device=local_slab.device)
    dist.all_gather_into_tensor(out, local_slab.contiguous(), group=group)
    return out


def unpack_results(slab, N, n, m, p):
    zs, ls = N * (n + m), p * (N - 1) * n
    a = slab.cpu().numpy() if hasattr(slab, "cpu") else np.asarray(slab)
    B = a.shape[0]
    return {"Z": a[:, :zs].reshape(B, N, n + m), "L": a[:, zs:zs + ls].reshape(B, p, N - 1, n),
            "stats": a[:, zs + ls:zs + ls + 10], "status": a[:, -1].astype(np.int32)}


def pack_host_results(out):
    """Same slab as pack_results, built from the host arrays agb_newton_solve_batch returned."""
    import torch
    B = out["Z"].shape[0]
    slab = np.concatenate([out["Z"].reshape(B, -1), out["L"].reshape(B, -1), out["stats"],
                           out["status"].astype(np.float64)[:, None]], axis=1)
    return torch.from_numpy(np.ascontiguousarray(slab))


def solve_sharded(make_batch, x0, xf, Z0, L0, opts, rank, world, gather=True):
    """Shard a batch contiguously across `world` ranks, solve the local shard, all-gather the results.
    `make_batch(local_B)` returns a GameBatch on this rank's device.  Shards may be uneven (total % world != 0): every rank
    pads its slab to the largest shard for the collective and the padding is trimmed afterwards."""
    total = x0.shape[0]
    if world > total:
        raise ValueError(f"cannot shard {total} instances over {world} ranks: every rank needs at least one instance")
    lo, hi = shard_bounds(total, world, rank)
    gb = make_batch(hi - lo)
    gb.set_instance_params(x0=x0[lo:hi], xf=None if xf is None else xf[lo:hi])
    gb.set_initial(Z0[lo:hi], L0[lo:hi])
    out = gb.newton_solve(opts)
    slab = pack_host_results(out)
    if gather and world > 1:
        import torch
        sizes = [shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world)]
        pad = max(sizes)
        if slab.shape[0] < pad:
            slab = torch.cat([slab, torch.zeros((pad - slab.shape[0], slab.shape[1]), dtype=slab.dtype)], dim=0)
        full = all_gather_results(slab).view(world, pad, -1)
        slab = torch.cat([full[r, :sizes[r]] for r in range(world)], dim=0)
    res = unpack_results(slab, gb.N, gb.n, gb.m, gb.p)
    gb.close()
    return res
