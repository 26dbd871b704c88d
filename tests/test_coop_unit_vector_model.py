"""CPU model of coop_unit_vector (raytracingincuda_b200/csrc/rt_kernels.cu): the warp-cooperative tail of random_unit_vector's
rejection loop (GF vec3.h:117-127).

The sampling spec is sequential: a lane tries block 0, 1, 2, ... of its own Philox counter and takes the first candidate that
lies in the unit ball.  The kernel lets every lane of the warp try one block per trip for the first searching lane at or above
it.  This file transcribes the kernel's ballots, shuffles and masks literally (one Python loop per warp-wide statement) and
checks, for random sets of searching lanes and random accept tables, that every searching lane ends with exactly the candidate
the sequential loop would have accepted, and that the number of trips is what makes it worth doing."""
import numpy as np
import pytest


def ffs(x):
    return (x & -x).bit_length()          # 1-based index of the lowest set bit, 0 for 0 (CUDA __ffs)


def clz(x):
    return 32 - x.bit_length()            # CUDA __clz on a 32-bit word


def coop(need, accept):
    """need[lane] -> bool; accept[lane][k] -> whether block k of that lane's counter yields a point in the ball (k >= 1).
    Returns (chosen block per lane or None, trips)."""
    need = list(need)
    nxt = [1] * 32
    chosen = [None] * 32
    searching = sum(1 << l for l in range(32) if need[l])
    trips = 0
    while searching:
        trips += 1
        up = [searching >> l for l in range(32)]
        dist = [ffs(up[l]) - 1 if up[l] else 0 for l in range(32)]
        f = [l + dist[l] for l in range(32)]
        hk = [nxt[f[l]] + dist[l] for l in range(32)]                       # __shfl_sync(next, f) + dist
        ok = [bool(up[l]) and accept[f[l]][hk[l]] for l in range(32)]       # the helper computes block hk of lane f's counter
        passed = sum(1 << l for l in range(32) if ok[l])
        src = list(range(32))
        found = [False] * 32
        for l in range(32):
            if need[l]:
                below = searching & ((1 << l) - 1)
                prev = 31 - clz(below) if below else -1
                upto_me = 0xffffffff >> (31 - l)
                upto_prev = (0xffffffff >> (31 - prev)) if prev >= 0 else 0
                mine = passed & upto_me & ~upto_prev & 0xffffffff
                if mine:
                    src[l] = 31 - clz(mine)
                    found[l] = True
                    need[l] = False
                else:
                    nxt[l] += l - prev
        for l in range(32):
            if found[l]:
                chosen[l] = hk[src[l]]                                        # the candidate comes from the helper's registers
                assert f[src[l]] == l                                         # ... and it was computed for this lane's counter
        searching = sum(1 << l for l in range(32) if need[l])
    return chosen, trips


@pytest.mark.parametrize("p_need", [1.0, 0.476, 0.1, 1 / 32])
def test_every_searching_lane_gets_the_sequential_loops_candidate(p_need):
    rng = np.random.default_rng(int(p_need * 1000))
    trips_total = seq_total = warps = 0
    for _ in range(300):
        need = rng.random(32) < p_need
        if p_need == 1 / 32:
            need[:] = False
            need[rng.integers(0, 32)] = True
        accept = rng.random((32, 400)) < np.pi / 6                            # the ball fills pi/6 of the cube
        accept[:, 399] = True
        chosen, trips = coop(need, accept)
        for l in range(32):
            if need[l]:
                first = 1 + int(np.argmax(accept[l, 1:]))
                assert chosen[l] == first, (l, chosen[l], first)
            else:
                assert chosen[l] is None
        if need.any():
            warps += 1
            trips_total += trips
            seq_total += max(1 + int(np.argmax(accept[l, 1:])) for l in range(32) if need[l])
    # the sequential loop makes the warp wait for its unluckiest lane; the cooperative one needs about half the trips
    assert trips_total <= seq_total
    if p_need >= 0.1:
        assert trips_total < 0.62 * seq_total, (trips_total / warps, seq_total / warps)


def test_adversarial_masks():
    accept = np.zeros((32, 128), dtype=bool)
    accept[:, 40] = True                                                      # every lane's first passing block is 40
    for mask in (0xffffffff, 0x80000000, 0x00000001, 0xaaaaaaaa, 0x55555555, 0x80000001, 0x0000ffff, 0xffff0000, 0x00010000):
        need = [(mask >> l) & 1 == 1 for l in range(32)]
        chosen, _ = coop(need, accept)
        assert all((chosen[l] == 40) == need[l] for l in range(32)), hex(mask)
