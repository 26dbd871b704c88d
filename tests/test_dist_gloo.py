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
        chunks = rt.num_chunks(W, H, SPP)

        def render_partials(planes, c0, c1):
            for c in range(c0, c1):
                for p in range(W * H):
                    acc = np.zeros(3, dtype=np.float32)
                    for s in range(c * SPP // chunks, (c + 1) * SPP // chunks):
                        acc = acc + O.sample(slots, cam, p % W, p // W, s)
                    planes[c - c0, p, :3] = torch.from_numpy(acc)

        def finalize(planes):
            acc = torch.zeros((W * H, 3), dtype=torch.float32)
            for c in range(planes.shape[0]):
                acc = acc + planes[c, :, :3]
            v = (acc * torch.tensor(cam.scale, dtype=torch.float32)).numpy()
            # numpy's sqrt is correctly rounded; torch's vectorised CPU sqrt is not
            g = np.where(v > 0, np.sqrt(v), np.float32(0)).astype(np.float32)
            return torch.from_numpy(g).reshape(H, W, 3)
        frame = rtdist.render_spp_split(render_partials, finalize, W, H, chunks, rank, world, dev,
                                        combine="reduce" if mode == "spp-reduce" else "gather")
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
                                             (2, "spp-reduce", 29614)])
def test_split_assembles_to_single_process_frame(world, mode, port):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    want, _ = O.render(O.scene(3), O.camera(W, H, SPP, DEPTH))
    got = _run(world, mode, port)
    assert got.shape == want.shape
    if mode == "spp-reduce":
        # a sum-reduce across ranks re-associates the chunk sums: last-ulp differences are allowed
        assert np.allclose(got, want, rtol=0, atol=2e-6)
    else:
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
