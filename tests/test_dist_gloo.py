"""Host-side multi-GPU logic on CPU: world_size-2 (and 3) gloo groups run the frame partitioning
and the rank-0 assembly of raytracingincuda_b200.dist with the CPU oracle injected as the
renderer.  The assembled frame must equal the single-process frame bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, SPP, DEPTH = 24, 19, 16, 10


def _worker(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    import raytracingincuda_b200 as rt
    from raytracingincuda_b200 import dist as rtdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    slots, cam = O.scene(3), O.camera(W, H, SPP, DEPTH)
    dev = torch.device("cpu")
    if mode == "rows":  # noqa: E501
        def render_rows(buf):
            rows = rt.partition_rows(H, 4, rank, world)
            for k, j in enumerate(rows):
                img, _ = O.render(slots, cam, row0=int(j), row1=int(j) + 1)
                buf[k] = torch.from_numpy(img[0])
        frame = rtdist.render_rows_split(render_rows, W, H, 4, rank, world, dev)
    else:
        def render_partials(acc, s0, s1):
            acc.copy_(torch.from_numpy(O.accumulate(slots, cam, s0, s1)))

        def finalize(acc):
            return torch.from_numpy(O.finalize(acc.numpy(), cam))
        frame = rtdist.render_spp_split(render_partials, finalize, W, H, SPP, rank, world, dev)
    if rank == 0:
        q.put(frame.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def _run(world, mode, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    frame = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return frame


@pytest.mark.parametrize("world,mode,port", [(2, "rows", 29611), (3, "rows", 29612), (2, "spp", 29613),
                                             (3, "spp", 29614)])
def test_split_assembles_to_single_process_frame(world, mode, port):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    want, _ = O.render(O.scene(3), O.camera(W, H, SPP, DEPTH))
    got = _run(world, mode, port)
    assert got.shape == want.shape
    # rows: gather; spp: ONE integer sum-reduce of the accumulation buffer -- both bit-identical to the single-process frame
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
