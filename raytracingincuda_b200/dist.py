"""Multi-GPU frame partitioning: one process per GPU over ``torch.distributed``.

The path shards trivially -- pixels and samples are independent and the scene (<= 24 KB, or a few MB with
its LBVH) is replicated -- so there is no data-path collective inside the render.  The only exchange is
the final assembly on rank 0 (BASELINE.json north_star):

  rows : interleaved tiles of ``tile_rows`` rows, tile t -> rank t mod world.  Each rank renders
         and gamma-encodes its own rows; rank 0 gathers them (NCCL gather over NVLink) and
         scatters them to their row positions.  Philox is keyed by the global pixel index, so the
         assembled frame is bit-identical to the 1-GPU frame.
  spp  : rank r renders samples [S*r/world, S*(r+1)/world) of every pixel into the int64 fixed-point
         accumulation buffer (W*H*3 words); ONE sum-reduce of that buffer (NCCL over NVLink/NVSwitch)
         delivers the total to rank 0, which scales, gamma-encodes and stores.  The sums are integers, so
         the reduce is exact in any association order and the frame is bit-identical to the 1-GPU frame.

The functions take a ``render``/``render_partials`` callable so the host logic can be exercised on
CPU with the gloo backend (tests/test_dist_gloo.py injects the oracle there).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import api


def _gather_to_rank0(local, rank, world, group=None):
    """Gather equally-shaped tensors to rank 0; returns the list on rank 0, None elsewhere."""
    if world == 1:
        return [local]
    bufs = [torch.empty_like(local) for _ in range(world)] if rank == 0 else None
    dist.gather(local, gather_list=bufs, dst=0, group=group)
    return bufs


def rows_of(height, tile_rows, rank, world):
    return api.partition_rows(height, tile_rows, rank, world)


def render_rows_split(render_rows, width, height, tile_rows, rank, world, device, dtype=torch.float32, group=None,
                      out=None):
    """``render_rows(buf)`` must fill ``buf`` ((n_local_rows, width, 3) on ``device``) with this
    rank's rows in ascending row order.  Returns the assembled (height, width, 3) frame on rank 0."""
    my_rows = rows_of(height, tile_rows, rank, world)
    max_rows = max(len(rows_of(height, tile_rows, r, world)) for r in range(world))
    local = torch.zeros((max_rows, width, 3), dtype=dtype, device=device)
    render_rows(local[:len(my_rows)])
    parts = _gather_to_rank0(local, rank, world, group)
    if rank != 0:
        return None
    frame = out if out is not None else torch.empty((height, width, 3), dtype=dtype, device=device)
    for r, part in enumerate(parts):
        rows = torch.from_numpy(rows_of(height, tile_rows, r, world).astype(np.int64)).to(device)
        frame.index_copy_(0, rows, part[:len(rows)])
    return frame


def render_spp_split(render_partials, finalize, width, height, spp, rank, world, device, group=None, acc=None):
    """``render_partials(acc, s0, s1)`` overwrites ``acc`` (a contiguous (height, width, 3) int64 tensor on
    ``device``; pass one to reuse it across frames) with this rank's fixed-point radiance sums over samples
    [s0, s1); ``finalize(acc)`` turns the summed accumulators into the frame on rank 0.  Returns the frame on
    rank 0, None elsewhere."""
    s0, s1 = api.partition_samples(spp, rank, world)
    if acc is None:
        acc = torch.empty((height, width, 3), dtype=torch.int64, device=device)
    if s1 > s0:
        render_partials(acc, s0, s1)
    else:
        acc.zero_()
    if world > 1:
        dist.reduce(acc, dst=0, op=dist.ReduceOp.SUM, group=group)      # integer sum: exact in any order
    return finalize(acc) if rank == 0 else None
