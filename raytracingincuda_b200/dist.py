"""Multi-GPU frame partitioning: one process per GPU over ``torch.distributed``.

The path shards trivially -- pixels and samples are independent and the scene (<= 24 KB) is
replicated -- so there is no data-path collective inside the render.  The only exchange is the
final assembly on rank 0 (BASELINE.json north_star):

  rows : interleaved tiles of ``tile_rows`` rows, tile t -> rank t mod world.  Each rank renders
         and gamma-encodes its own rows; rank 0 gathers them (NCCL gather over NVLink) and
         scatters them to their row positions.  Philox is keyed by the global pixel index, so the
         assembled frame is bit-identical to the 1-GPU frame.
  spp  : rank r renders chunks [C*r/world, C*(r+1)/world) of every pixel into linear float4
         planes; rank 0 gathers the planes and ``rt_finalize`` adds them in chunk order -- the same
         order the 1-GPU path uses, so this split is bit-identical too (a plain NCCL sum-reduce
         would not be: its association order is not defined).

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


def render_spp_split(render_partials, finalize, width, height, chunks, rank, world, device, group=None,
                     combine="gather"):
    """``render_partials(planes, c0, c1)`` fills ``planes`` ((c1-c0, width*height, 4) float32 on
    ``device``) with this rank's chunk sums; ``finalize(all_planes)`` turns the (chunks, W*H, 4)
    stack into the frame on rank 0.

    combine="gather" (default): rank 0 receives every plane and adds them in chunk order --
    bit-identical to the 1-GPU frame.  combine="reduce": every rank adds its own planes in chunk
    order and one NCCL sum-reduce of the W*H*4 accumulation buffer delivers the total to rank 0 --
    1/world of the traffic, but the association order of the cross-rank sum is NCCL's, so the
    frame can differ from the canonical one in the last ulp."""
    bounds = [api.partition_chunks(chunks, r, world) for r in range(world)]
    c0, c1 = bounds[rank]
    max_c = max(b[1] - b[0] for b in bounds)
    local = torch.zeros((max_c, width * height, 4), dtype=torch.float32, device=device)
    if c1 > c0:
        render_partials(local[:c1 - c0], c0, c1)
    if combine == "reduce":
        acc = local[0].clone()
        for k in range(1, c1 - c0):
            acc += local[k]
        if world > 1:
            dist.reduce(acc, dst=0, op=dist.ReduceOp.SUM, group=group)
        return finalize(acc.unsqueeze(0).contiguous()) if rank == 0 else None
    parts = _gather_to_rank0(local, rank, world, group)
    if rank != 0:
        return None
    if world == 1:
        planes = local[:chunks]
    else:
        planes = torch.cat([p[:b[1] - b[0]] for p, b in zip(parts, bounds)], dim=0)
    return finalize(planes.contiguous())

