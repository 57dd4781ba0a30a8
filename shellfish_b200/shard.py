"""Multi-GPU sharding of the halo list (one process per GPU).

Halos are independent (each los.Halo only reads particles, cmd/shell.go:406-408),
so the halo table is partitioned across ranks by estimated particle count and
every rank runs the unchanged single-GPU path on its share: no data-path
collective.  The only exchange is gathering the per-halo coefficient rows back
into catalog order on rank 0, through torch.distributed (NCCL on GPUs, gloo in
the CPU tests).
"""
import numpy as np


def partition_halos(weights, n_ranks, groups=None):
    """Longest-processing-time greedy partition.

    weights: per-halo cost estimate (particle count within the LoS sphere, or
    R200m**3 as a proxy before the first pass).  groups: optional integer label
    per halo (the snapshot block holding its centre, `block_of`); halos of one
    group go to the same rank, so a rank reads few particle blocks -- it only
    needs the blocks its own halos touch (`blocks_needed`).  Returns a list of
    n_ranks index arrays; each keeps catalog order so that Transform's
    cumulative in-place wrap (los/halo.go:168-197) sees the rank's halos in the
    same relative order as the reference loop does.
    """
    weights = np.asarray(weights, np.float64)
    if groups is None:
        groups = np.arange(len(weights))
    groups = np.asarray(groups, np.int64)
    labels, inv = np.unique(groups, return_inverse=True)
    gw = np.bincount(inv, weights=weights, minlength=len(labels))
    order = np.argsort(-gw, kind="stable")
    loads = np.zeros(n_ranks)
    owner = np.zeros(len(labels), np.int64)
    for g in order:
        r = int(np.argmin(loads))
        owner[g] = r
        loads[r] += gw[g]
    halo_owner = owner[inv]
    return [np.flatnonzero(halo_owner == r).astype(np.int64) for r in range(n_ranks)]


def morton_keys(xyz, total_width, bits=10):
    """Z-curve keys of positions in [0, total_width)^3, `bits` bits per axis (x most significant)."""
    xyz = np.asarray(xyz, np.float64)
    q = np.clip(np.floor(xyz / float(total_width) * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    key = np.zeros(len(q), np.int64)
    for b in range(bits):
        for ax in range(3):
            key |= ((q[:, ax] >> b) & 1) << (3 * b + (2 - ax))
    return key


def partition_spatial(weights, xyz, total_width, n_ranks):
    """Partition by estimated particle count along a space-filling curve: halos sorted by
    the Z-curve key of their centre, the sequence cut into n_ranks contiguous runs of equal
    weight (a halo goes to the run holding the midpoint of its weight interval).  Each rank's
    halos are spatially compact, so it touches few particle blocks (`blocks_needed`): for the
    1000-halo configs[1] box on 8 ranks ~120 of 512 blocks per rank against ~200 for the plain
    LPT of `partition_halos`, at 0.5% against 0.3% load imbalance.  Returns index arrays in
    catalog order, like partition_halos."""
    weights = np.asarray(weights, np.float64)
    if len(weights) == 0:
        return [np.zeros(0, np.int64) for _ in range(n_ranks)]
    order = np.argsort(morton_keys(xyz, total_width), kind="stable")
    wv = weights[order]
    cum = np.cumsum(wv)
    total = cum[-1] if cum[-1] > 0 else 1.0
    own = np.minimum(((cum - wv / 2) / total * n_ranks).astype(np.int64), n_ranks - 1)
    return [np.sort(order[own == r]).astype(np.int64) for r in range(n_ranks)]


def block_of(halos, blocks):
    """Index of the block whose [Origin, Origin + Width) holds each halo centre
    (-1 if none does): the affinity label for partition_halos."""
    xyz = np.asarray(halos, np.float64)[:, :3]
    out = np.full(len(xyz), -1, np.int64)
    for bi, b in enumerate(blocks):
        o = np.asarray(b["Origin"], np.float64)
        w = np.asarray(b["Width"], np.float64)
        m = np.all((xyz >= o) & (xyz < o + w), axis=1) & (out < 0)
        out[m] = bi
    return out


def blocks_needed(ctx, blocks, owned):
    """Indices of the blocks this rank has to push: those a halo it owns touches
    (binIntersections, cmd/shell.go:599-611, restricted to the rank's halos).
    A pushed block still replays Transform for every halo touching it, owned or
    not (sfk_set_owned), so results do not depend on the partition."""
    owned = np.asarray(owned, bool)
    need = []
    for bi, b in enumerate(blocks):
        hit = ctx.block_halo_mask(b["Origin"], b["Width"], b["TotalWidth"])
        if np.any(hit & owned):
            need.append(bi)
    return need


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


def gather_table(parts, rank, local_rows, n_total, dist=None, device="cpu"):
    """gather_rows when every rank knows the whole partition (`parts` from
    partition_halos): one fixed-size all-gather, no size or index exchange.
    local_rows: (len(parts[rank]), k).  Returns the (n_total, k) table."""
    local_rows = np.asarray(local_rows, np.float64)
    k = local_rows.shape[1]
    table = np.zeros((n_total, k))
    world = len(parts)
    if world == 1 or dist is None or not dist.is_initialized():
        table[parts[rank]] = local_rows
        return table
    import torch
    n_max = max(1, max(len(p) for p in parts))
    loc = torch.zeros((n_max, k), dtype=torch.float64)
    loc[:len(local_rows)] = torch.from_numpy(local_rows)
    loc = loc.to(device)
    full = torch.empty((world * n_max, k), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(full, loc)
    full = full.cpu().numpy().reshape(world, n_max, k)
    for r, p in enumerate(parts):
        table[p] = full[r, :len(p)]
    return table


def comm_init(ctx, dist, device):
    """Joins `ctx` to an NCCL communicator owned by libsfk (sfk_comm_init): rank 0 draws
    the unique id, torch.distributed carries its 128 bytes to the other ranks."""
    import torch
    from . import api
    rank, world = dist.get_rank(), dist.get_world_size()
    raw = api.comm_unique_id() if rank == 0 else bytes(api.COMM_ID_BYTES)
    t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
    dist.broadcast(t, src=0)
    ctx.comm_init(bytes(t.cpu().numpy().tobytes()), rank, world)


def split_halo_finish(ctx, dist, in_library=False):
    """Split-halo exchange (BASELINE config 3): every rank has pushed its share
    of the particle blocks into `ctx`.  in_library: the whole exchange runs inside
    libsfk over its own NCCL communicator (comm_init first): reduce-scatter of the
    grid by ring, ring-partitioned analysis, all-gather of the filtered points.
    Otherwise the three all-reduces below, issued through torch.distributed on the
    library's exchange buffers, replace Halo.Split/Join (los/halo.go:107-140) and
    every rank analyses every ring.  Returns (coeffs, status), identical on every
    rank."""
    from . import api
    if in_library:
        return ctx.split_finish_comm()
    ctx.split_count()
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(ctx.split_buffer(api.SPLIT_HIT_COUNTS), op=dist.ReduceOp.SUM)
        dist.all_reduce(ctx.split_buffer(api.SPLIT_RHO_MAX), op=dist.ReduceOp.MAX)
    ctx.split_bin()
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(ctx.split_buffer(api.SPLIT_GRID), op=dist.ReduceOp.SUM)
    return ctx.split_finish()
