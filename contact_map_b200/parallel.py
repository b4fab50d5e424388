"""Frame sharding across GPUs: one process per GPU, ``torch.distributed`` for the plumbing.

The reference's only parallel strategy is data parallelism over frames with a Python-level
reduce (``dask_run``: /root/reference/contact_map/dask_runner.py:11-37; partition
``frequency_task.default_slices`` :49-67; reduce ``reduce_all_results`` :125-141, i.e.
``Counter`` addition).  Here every rank owns the frame blocks ``k`` with ``k % world == rank``,
accumulates them into count tables of IDENTICAL layout, and the partial tables are summed
with ONE all-reduce (NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU tests).  Integer
addition is associative, so the result is independent of the partition.
"""
import numpy as np
import torch
import torch.distributed as dist

from .frequency_task import default_slices


def world():
    """(rank, world_size) of the default process group, (0, 1) when not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def rank_slices(n_frames, n_blocks=None, rank=None, world_size=None):
    """Frame slices of this rank: the reference partition ``default_slices(n_frames, n_blocks)``
    dealt round-robin to the ranks.  ``n_blocks`` defaults to the world size."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    slices = default_slices(n_frames, n_blocks or world_size)
    return [s for k, s in enumerate(slices) if k % world_size == rank]


def reduce_tables(tables, group=None, dst=None):
    """Sum count tables over all ranks, in place, with a single collective.

    ``tables`` is a list of int32 tensors (views of the uint32 tables).  They are reduced as one
    flat buffer so exactly one collective is issued; two's-complement addition makes the int32
    view exact for uint32 counts.  ``dst=None``: all-reduce, every rank ends up with the sum.
    ``dst=r``: reduce to rank ``r`` only -- what the reference's ``reduce_all_results`` does (the
    client alone receives the result, frequency_task.py:125-141) at about half the traffic; the
    tables of the other ranks are unspecified afterwards."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tables

    def collective(flat):
        if dst is None:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.reduce(flat, dst=dst, op=dist.ReduceOp.SUM, group=group)
    hashed = [t for t in tables if hasattr(t, "slots")]
    if hashed:
        # hashed tables have rank-dependent layouts: gather the sparse exports and merge them
        # (one all-gather per table; the merged (index, count) arrays become the table's export)
        from .engine import get_engine
        engine = get_engine(hashed[0].slots.device)
        for t, (idx, cnt) in zip(hashed, engine.export_many(hashed)):
            t.merged = merge_sparse(idx, cnt, t.slots.device, group)
        tables = [t for t in tables if not hasattr(t, "slots")]
        if not tables:
            return hashed
    if len(tables) > 1 and all(t.is_contiguous() and t.dtype == tables[0].dtype and
                               t.untyped_storage().data_ptr() == tables[0].untyped_storage().data_ptr()
                               for t in tables):
        # views of one allocation (engine.new_tables): reduce the span they cover in place
        lo = min(t.storage_offset() for t in tables)
        hi = max(t.storage_offset() + t.numel() for t in tables)
        if hi - lo <= 2 * sum(t.numel() for t in tables):
            span = torch.empty(0, dtype=tables[0].dtype, device=tables[0].device)
            span.set_(tables[0].untyped_storage(), lo, (hi - lo,))
            collective(span)
            return tables
    flat = torch.cat([t.reshape(-1) for t in tables]) if len(tables) > 1 else tables[0].reshape(-1)
    collective(flat)
    if len(tables) > 1:
        offset = 0
        for t in tables:
            t.copy_(flat[offset:offset + t.numel()].reshape(t.shape))
            offset += t.numel()
    return tables


# dense tables above this many bytes are not all-reduced: their sparse exports are merged instead
# (the 100 000-atom system has a 40 GB atom table holding ~3 M non-zero entries)
SPARSE_REDUCE_BYTES = 1 << 30


def reduce_export(engine, tables, dst=None, group=None):
    """Sum the count tables over all ranks AND export them: the multi-GPU replacement of the
    reference's ``reduce_all_results`` (frequency_task.py:125-141).  Returns
    ``[(flat index int64[], count uint32[]), ...]`` sorted by index on the ranks that receive the
    result (every rank for ``dst=None``), ``None`` on the others.

    Small dense tables: ONE in-place collective over the tables (NCCL over NVLink), then the
    sparse export on the result rank.  Large dense tables and hashed tables: every rank exports
    its own non-zero entries (a few tens of MB) and the sparse lists are all-gathered and merged
    by key -- moving 40 GB per rank through the switch to add up 3 M numbers would cost more than
    the frames themselves."""
    rank, world_size = world()
    if world_size == 1:
        return engine.export_many(tables)
    dense_bytes = sum(t.numel() * 4 for t in tables if not hasattr(t, "slots"))
    if any(hasattr(t, "slots") for t in tables) or dense_bytes > SPARSE_REDUCE_BYTES:
        receives = dst is None or rank == dst
        packed = engine.export_packed_device(tables)
        merged = []
        for t, p in zip(tables, packed):
            if p is None:
                # hashed tables / counts beyond 24 bits: through the host lists
                (idx, cnt), = engine.export_many([t])
                device = t.slots.device if hasattr(t, "slots") else t.device
                m = merge_sparse(idx, cnt, device, group)
                merged.append(m if receives else None)
            else:
                merged.append(merge_packed(p, receives, group))
        return merged if receives else None
    reduce_tables(tables, group=group, dst=dst)
    return engine.export_many(tables) if dst is None or rank == dst else None


def merge_packed(packed, receives=True, group=None):
    """Sum the ranks' sparse exports WITHOUT leaving the device: ``packed`` is this rank's sorted
    int64 tensor of (flat index << 24 | count) words (``engine.export_packed_device``); the lists
    are all-gathered (NCCL over NVLink: a few tens of MB per rank), sorted by index and summed per
    index; only the ranks that receive the result copy it to the host.
    Returns (index int64[], count uint32[]) sorted by index, or ``None``."""
    w = dist.get_world_size(group)
    device = packed.device
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(w)]
    dist.all_gather(sizes, torch.tensor([packed.numel()], dtype=torch.int64, device=device), group=group)
    sizes = [int(x.item()) for x in sizes]
    cap = max(max(sizes), 1)
    mine = torch.zeros(cap, dtype=torch.int64, device=device)
    mine[:packed.numel()] = packed
    parts = [torch.empty(cap, dtype=torch.int64, device=device) for _ in range(w)]
    dist.all_gather(parts, mine, group=group)
    if not receives:
        return None
    everything = torch.cat([p[:k] for p, k in zip(parts, sizes)])
    uniq, inverse = torch.unique(everything >> 24, sorted=True, return_inverse=True)
    total = torch.zeros(uniq.numel(), dtype=torch.int64, device=device).scatter_add_(0, inverse, everything & 0xffffff)
    return uniq.cpu().numpy(), total.cpu().numpy().astype(np.uint32)


def merge_sparse(idx, cnt, device, group=None):
    """Sum sparse (flat index, count) lists over all ranks: all-gather + sort + segment sum.
    Returns (index int64[], count uint32[]) sorted by index, identical on every rank."""
    w = dist.get_world_size(group)
    li = torch.as_tensor(np.ascontiguousarray(idx, dtype=np.int64), device=device)
    lc = torch.as_tensor(np.ascontiguousarray(cnt).astype(np.int64), device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(w)]
    dist.all_gather(sizes, torch.tensor([li.numel()], dtype=torch.int64, device=device), group=group)
    sizes = [int(x.item()) for x in sizes]
    cap = max(max(sizes), 1)
    packed = torch.zeros((2, cap), dtype=torch.int64, device=device)
    packed[0, :li.numel()] = li
    packed[1, :lc.numel()] = lc
    parts = [torch.empty((2, cap), dtype=torch.int64, device=device) for _ in range(w)]
    dist.all_gather(parts, packed, group=group)
    all_idx = torch.cat([p[0, :k] for p, k in zip(parts, sizes)])
    all_cnt = torch.cat([p[1, :k] for p, k in zip(parts, sizes)])
    uniq, inverse = torch.unique(all_idx, sorted=True, return_inverse=True)
    total = torch.zeros(uniq.numel(), dtype=torch.int64, device=device).scatter_add_(0, inverse, all_cnt)
    return uniq.cpu().numpy(), total.cpu().numpy().astype(np.uint32)


def reduce_frame_count(n_frames, device, group=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return int(n_frames)
    t = torch.tensor([int(n_frames)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def gather_keys(keys, device, group=None):
    """Concatenate per-rank uint64 key arrays on every rank (ContactTrajectory: no reduction,
    rank-ordered concatenation; the keys carry global frame numbers, so a sort restores
    frame order)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return keys
    w = dist.get_world_size(group)
    local = torch.as_tensor(keys.view(np.int64), device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(w)]
    dist.all_gather(sizes, torch.tensor([local.numel()], dtype=torch.int64, device=device), group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    padded = torch.zeros(cap, dtype=torch.int64, device=device)
    padded[:local.numel()] = local
    parts = [torch.empty(cap, dtype=torch.int64, device=device) for _ in range(w)]
    dist.all_gather(parts, padded, group=group)
    merged = torch.cat([p[:s] for p, s in zip(parts, sizes)]).cpu().numpy().view(np.uint64)
    return np.sort(merged)
