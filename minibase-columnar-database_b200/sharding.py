"""TID-range sharding of a scan across ranks: one process per GPU, rows split by contiguous
position range, no data-path collective.  After the local scans, torch.distributed (NCCL on GPUs, gloo in the CPU
tests) only moves results: every rank's [aggregates, count] block is all-gathered and folded locally (= the
all-reduce of COUNT/SUM/MIN/MAX), and position / value lists are gathered on rank 0 at the exclusive-scan offsets
of the counts -- rank order is position order, so the concatenation is already sorted."""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.distributed as dist

TILE = 8192        # shard boundaries are multiples of the column padding (and of the 8000-bit bitmap page grain x tile)


def shard_range(total_rows: int, world: int, rank: int) -> tuple[int, int]:
    """[begin, end) of `rank`: equal tile-aligned slices, the last rank takes the ragged tail."""
    per = (total_rows + world - 1) // world
    per = (per + TILE - 1) // TILE * TILE
    lo = min(total_rows, rank * per)
    hi = min(total_rows, lo + per) if rank < world - 1 else total_rows
    return lo, max(lo, hi)


def fold_aggregates(blocks: torch.Tensor, kinds: Sequence[int], real: Sequence[bool]) -> list:
    """blocks: [world, nagg+1] int64 raw 8-byte aggregate values (doubles as bits) + count in the last column.
    Returns python values [(value, valid)] per aggregate and the total count."""
    world = blocks.shape[0]
    counts = blocks[:, -1]
    total = int(counts.sum())
    out = []
    for a, (kind, is_real) in enumerate(zip(kinds, real)):
        col = blocks[:, a]
        vals = col.view(torch.float64) if is_real else col
        nonempty = counts > 0
        if kind in (0, 1):                                   # COUNT, SUM
            v = vals.sum()
            out.append((float(v) if is_real else int(v), True))
        elif not bool(nonempty.any()):
            out.append((0.0 if is_real else 0, False))        # MIN/MAX over an empty set
        else:
            sel = vals[nonempty]
            v = sel.min() if kind == 2 else sel.max()
            out.append((float(v) if is_real else int(v), True))
    return out, total


def allgather_blocks(block: torch.Tensor, group=None) -> torch.Tensor:
    """Every rank's 1-D block -> [world, len(block)] on every rank (one collective)."""
    world = dist.get_world_size(group)
    if block.is_cuda:                                        # NCCL: one call, no per-rank tensors
        out = torch.empty(world * block.numel(), dtype=block.dtype, device=block.device)
        dist.all_gather_into_tensor(out, block, group=group)
        return out.view(world, block.numel())
    parts = [torch.empty_like(block) for _ in range(world)]
    dist.all_gather(parts, block, group=group)
    return torch.stack(parts)


def gather_rows_multi(locals_: Sequence[tuple], counts: Sequence[int], dst: int = 0, group=None, wait: bool = True):
    """Gather several variable-length byte buffers at once (positions + every projected column) on `dst` in rank
    order with ONE batch of point-to-point operations.  locals_ = [(1-D uint8 tensor of counts[rank]*row_bytes bytes,
    row_bytes)].  Returns the concatenations on dst, None elsewhere; with wait=False returns (work handles, concatenations)
    and the caller waits (the transfers then overlap whatever is launched next)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ops, totals = [], []
    for local, row_bytes in locals_:
        if rank == dst:
            total = torch.empty(sum(counts) * row_bytes, dtype=torch.uint8, device=local.device)
            offs = [0]
            for c in counts:
                offs.append(offs[-1] + c * row_bytes)
            total[offs[rank]:offs[rank + 1]].copy_(local)
            ops += [dist.P2POp(dist.irecv, total[offs[r]:offs[r + 1]], r, group) for r in range(world) if r != dst and counts[r] > 0]
            totals.append(total)
        elif counts[rank] > 0:
            ops.append(dist.P2POp(dist.isend, local, dst, group))
    works = dist.batch_isend_irecv(ops) if ops else []
    if not wait:
        return works, (totals if rank == dst else None)
    for w in works:
        w.wait()
    return totals if rank == dst else None


def gather_rows_packed(locals_: Sequence[tuple], counts: Sequence[int], dst: int = 0, group=None):
    """Same result as gather_rows_multi with ONE message per rank: every rank packs its buffers back to back
    ([buffer 0 rows | buffer 1 rows | ...], one device copy), dst receives world - 1 blocks and splits them back into
    the per-buffer concatenations in rank order.  With many ranks the fixed cost of a point-to-point operation
    (not its bytes) is what a result gather pays, so fewer, larger messages win."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    widths = [rb for _, rb in locals_]
    row_total = sum(widths)
    mine = torch.cat([t for t, _ in locals_]) if len(locals_) > 1 else locals_[0][0]
    if rank != dst:
        if counts[rank] > 0:
            for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, mine, dst, group)]):
                w.wait()
        return None
    blocks = [mine if r == rank else torch.empty(counts[r] * row_total, dtype=torch.uint8, device=mine.device) for r in range(world)]
    ops = [dist.P2POp(dist.irecv, blocks[r], r, group) for r in range(world) if r != dst and counts[r] > 0]
    for w in (dist.batch_isend_irecv(ops) if ops else []):
        w.wait()
    outs, before = [], 0
    for rb in widths:
        outs.append(torch.cat([blocks[r][counts[r] * before:counts[r] * (before + rb)] for r in range(world)]))
        before += rb
    return outs


def allgather_rows(locals_: Sequence[tuple], counts: Sequence[int], dst: Optional[int] = 0, group=None):
    """The same concatenations as gather_rows_multi through ONE all-gather: every rank packs its buffers into a block
    padded to the largest count (buffer j at byte max_count * sum(row_bytes[:j])), the blocks are all-gathered, and dst
    (every rank when dst is None) cuts the rank-ordered pieces back out.  NCCL's all-gather runs at NVSwitch collective
    bandwidth, its grouped send/recv at a fraction of it, so on 8 GPUs moving 8x the bytes this way is still ~3x faster
    than the many-to-one gather (measured, profiles/README.md)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    widths = [rb for _, rb in locals_]
    maxc = max(max(counts), 1)
    dev = locals_[0][0].device
    block = torch.empty(maxc * sum(widths), dtype=torch.uint8, device=dev)
    before = 0
    for (t, rb) in locals_:
        block[maxc * before:maxc * before + counts[rank] * rb].copy_(t)
        before += rb
    blocks = allgather_blocks(block, group)                  # [world, maxc * row_total]
    if dst is not None and rank != dst:
        return None
    outs, before = [], 0
    for rb in widths:
        outs.append(torch.cat([blocks[r, maxc * before:maxc * before + counts[r] * rb] for r in range(world)]))
        before += rb
    return outs


def gather_rows(local: torch.Tensor, counts: Sequence[int], row_bytes: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Single-buffer form of gather_rows_multi."""
    out = gather_rows_multi([(local, row_bytes)], counts, dst, group)
    return out[0] if out is not None else None


class PeerGather:
    """Result gather over NVLink PEER MEMORY: `dst` exports one receive buffer through CUDA IPC when the object is built;
    after that every rank copies its pieces straight into that buffer at the exclusive-scan offsets (rank order =
    position order).  The copies are device-to-device DMA on the copy engines -- no NCCL data movement, no SMs, so they
    overlap the scan kernels that occupy every SM -- and one tiny all-reduce, ordered after each rank's copies on its
    stream, tells `dst` that everything has landed.

    Measured on this pool's 2-GPU box (bench.py, MBC_BENCH_GATHER=peer): correct, but the cross-process copies ran at
    ~5 GB/s (10.3 ms per step against 2.4 ms with the NCCL all-gather) -- torch's copy into an IPC-mapped tensor of
    another process was staged through the host there instead of going over NVLink -- so bench.py does not use it by
    default.  Kept as the host-side half of the planned peer-memory gather.

    Build it collectively (every rank of the group); `gather()` is collective too.  Raises on every rank alike if the
    buffer cannot be shared, so the caller can fall back to `allgather_rows`."""

    def __init__(self, capacity_bytes: int, device: torch.device, dst: int = 0, group=None):
        self.group, self.dst = group, dst
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.capacity = int(capacity_bytes)
        self.device = device
        ok, payload, err = 1, [None], None
        try:
            if self.rank == dst:
                from torch.multiprocessing.reductions import reduce_tensor
                self.buf = torch.empty(self.capacity, dtype=torch.uint8, device=device)
                payload = [reduce_tensor(self.buf)]                 # (rebuild function, IPC handle + geometry)
        except Exception as e:                                    # noqa: BLE001 -- agreed on below
            ok, err = 0, e
        dist.broadcast_object_list(payload, src=dst, group=group)
        try:
            if self.rank != dst:
                if payload[0] is None:
                    raise RuntimeError("the destination rank could not export its buffer")
                fn, args = payload[0]
                self.remote = fn(*args)                             # a tensor on dst's device aliasing dst's buffer
                self.remote[:1].copy_(torch.zeros(1, dtype=torch.uint8, device=device))   # peer access works?
            else:
                self.remote = self.buf
        except Exception as e:                                    # noqa: BLE001
            ok, err = 0, e
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            raise RuntimeError(f"PeerGather unavailable: {err}")
        self.landed = torch.zeros(1, dtype=torch.int32, device=device)

    def gather(self, locals_: Sequence[tuple], counts: Sequence[int]):
        """locals_ = [(1-D uint8 tensor of counts[rank] * row_bytes bytes, row_bytes)].  On dst returns views of the
        receive buffer, one per input buffer, holding the concatenation in rank order; None elsewhere.  The views are
        overwritten by the next gather."""
        total = sum(counts)
        widths = [rb for _, rb in locals_]
        if total * sum(widths) > self.capacity:
            raise RuntimeError(f"PeerGather buffer of {self.capacity} bytes is too small for {total} rows of {sum(widths)} bytes")
        before_rows = sum(counts[:self.rank])
        base = 0
        for t, rb in locals_:
            if counts[self.rank]:
                self.remote[base + before_rows * rb:base + (before_rows + counts[self.rank]) * rb].copy_(t, non_blocking=True)
            base += total * rb
        dist.all_reduce(self.landed, group=self.group)            # after this, on dst's stream, every piece is in place
        if self.rank != self.dst:
            return None
        outs, base = [], 0
        for rb in widths:
            outs.append(self.buf[base:base + total * rb])
            base += total * rb
        return outs
