"""Multi-GPU sharding of the halo list (one process per GPU).

Halos are independent (each los.Halo only reads particles, cmd/shell.go:406-408),
so the halo table is partitioned across ranks by estimated particle count and
every rank runs the unchanged single-GPU path on its share: no data-path
collective.  The only exchange is gathering the per-halo coefficient rows back
into catalog order on rank 0, through torch.distributed (NCCL on GPUs, gloo in
the CPU tests).
"""
import numpy as np


def partition_halos(weights, n_ranks):
    """Longest-processing-time greedy partition.

    weights: per-halo cost estimate (particle count within the LoS sphere, or
    R200m**3 as a proxy before the first pass).  Returns a list of n_ranks
    index arrays; each keeps catalog order so that Transform's cumulative
    in-place wrap (los/halo.go:168-197) sees the rank's halos in the same
    relative order as the reference loop does.
    """
    weights = np.asarray(weights, np.float64)
    order = np.argsort(-weights, kind="stable")
    loads = np.zeros(n_ranks)
    parts = [[] for _ in range(n_ranks)]
    for i in order:
        r = int(np.argmin(loads))
        parts[r].append(int(i))
        loads[r] += weights[i]
    return [np.array(sorted(p), np.int64) for p in parts]


def estimate_weights(halos, r_max_mult=3.0, r_kernel_mult=0.2):
    """Cost proxy before any particle is read: volume of the culling sphere,
    (RMaxMult + RKernelMult)^3 * R200m^3 (los/halo.go:148-152)."""
    r = np.asarray(halos, np.float64)[:, 3]
    return np.where(r > 0, ((r_max_mult + r_kernel_mult) * r) ** 3, 0.0)


def gather_rows(local_idx, local_rows, n_total, dist=None, device="cpu"):
    """All-gathers per-halo result rows into catalog order.

    local_idx: this rank's halo indices; local_rows: (len(local_idx), k) array.
    With dist=None (single process) this is a scatter into the full table.
    """
    local_rows = np.asarray(local_rows, np.float64)
    k = local_rows.shape[1] if local_rows.ndim == 2 else 1
    full = np.zeros((n_total, k))
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        full[np.asarray(local_idx, np.int64)] = local_rows.reshape(len(local_idx), k)
        return full
    import torch
    world = dist.get_world_size()
    # fixed-size exchange: pad every rank's share to the largest one
    n_loc = torch.tensor([len(local_idx)], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n_loc)
    n_max = int(max(int(s.item()) for s in sizes))
    idx = torch.full((n_max,), -1, dtype=torch.int64, device=device)
    rows = torch.zeros((n_max, k), dtype=torch.float64, device=device)
    if len(local_idx):
        idx[:len(local_idx)] = torch.as_tensor(np.asarray(local_idx, np.int64), device=device)
        rows[:len(local_idx)] = torch.as_tensor(local_rows.reshape(len(local_idx), k), device=device)
    all_idx = [torch.empty_like(idx) for _ in range(world)]
    all_rows = [torch.empty_like(rows) for _ in range(world)]
    dist.all_gather(all_idx, idx)
    dist.all_gather(all_rows, rows)
    for i_t, r_t in zip(all_idx, all_rows):
        i_np, r_np = i_t.cpu().numpy(), r_t.cpu().numpy()
        m = i_np >= 0
        full[i_np[m]] = r_np[m]
    return full


def split_halo_finish(ctx, dist):
    """Split-halo exchange (BASELINE config 3): every rank has pushed its share
    of the particle blocks into `ctx`; the three all-reduces below replace
    Halo.Split/Join (los/halo.go:107-140).  Returns (coeffs, status), identical
    on every rank."""
    from . import api
    ctx.split_count()
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(ctx.split_buffer(api.SPLIT_HIT_COUNTS), op=dist.ReduceOp.SUM)
        dist.all_reduce(ctx.split_buffer(api.SPLIT_RHO_MAX), op=dist.ReduceOp.MAX)
    ctx.split_bin()
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(ctx.split_buffer(api.SPLIT_GRID), op=dist.ReduceOp.SUM)
    return ctx.split_finish()
