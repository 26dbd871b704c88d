"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the reference goldens.

Bars: scene bytes, primary (slot id, t) and the rendered float image are BIT-EXACT against the
oracle on the same seeded inputs; converged radiance is within MAE <= 1/255 per channel and
PSNR >= 40 dB of the reference's own global-float render (BASELINE.json north_star).
"""
import os

import numpy as np
import pytest

import oracle_lib as O
import raytracingincuda_b200 as rt
from raytracingincuda_b200 import api

pytestmark = pytest.mark.gpu


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def image_metrics(a, b):
    """Per-channel MAE (code values) and PSNR (dB) between two uint8 images."""
    d = a.astype(np.float64) - b.astype(np.float64)
    mae = np.abs(d).mean(axis=(0, 1))
    mse = (d ** 2).mean()
    psnr = 10 * np.log10(255.0 ** 2 / mse) if mse > 0 else np.inf
    return mae, psnr


@pytest.mark.parametrize("scene_id", [1, 2, 3])
@pytest.mark.parametrize("size", [(320, 192), (97, 61)])
def test_primary_hits_bit_exact_vs_oracle(renderer, scene_id, size):
    w, h = size
    slots = rt.scene(scene_id)
    renderer.upload_scene(slots)
    cam = rt.camera(w, h)
    ids, t = renderer.primary_hits(cam)
    oids, ot = O.primary(O.scene(scene_id), O.camera(w, h))
    assert np.array_equal(ids, oids)
    assert np.array_equal(bits(t), bits(ot))
    assert (ids >= 0).any() and (ids < 0).any()


@pytest.mark.parametrize("scene_id", [1, 2, 3])
def test_primary_hits_bit_exact_vs_oracle_double(renderer, scene_id):
    slots = rt.scene(scene_id, double=True)
    renderer.upload_scene(slots)
    cam = rt.camera(160, 96, double=True)
    ids, t = renderer.primary_hits(cam)
    oids, ot = O.primary(O.scene(scene_id, True), O.camera(160, 96, double=True))
    assert np.array_equal(ids, oids)
    assert np.array_equal(bits(t), bits(ot))


@pytest.mark.parametrize("scene_id", [1, 2, 3])
@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_primary_hits_bit_exact_vs_reference_golden(renderer, golden_dir, scene_id, tag):
    """(slot id, t) of the reference's own hit_world(), captured on a B200 (make_goldens_gpu.py)."""
    path = os.path.join(golden_dir, f"primary_scene{scene_id}_{tag}.npz")
    if not os.path.exists(path):
        pytest.skip("reference golden not generated yet")
    g = np.load(path)
    h, w = g["ids"].shape
    double = tag == "f64"
    renderer.upload_scene(rt.scene(scene_id, double=double))
    ids, t = renderer.primary_hits(rt.camera(w, h, double=double))
    assert np.array_equal(ids, g["ids"].astype(np.int32))
    assert np.array_equal(bits(t), bits(g["t"]))


@pytest.mark.parametrize("accel", ["linear", "lbvh", "grid", "auto"])
@pytest.mark.parametrize("scene_id,w,h,spp,depth", [(1, 48, 32, 12, 25), (2, 64, 40, 16, 50), (3, 40, 24, 9, 50),
                                                     (1, 33, 17, 3, 4)])
def test_render_bit_exact_vs_oracle(renderer, scene_id, w, h, spp, depth, accel):
    """Whole frames, bit for bit, against the oracle -- through every structure hit_world can use."""
    slots = rt.scene(scene_id)
    renderer.upload_scene(slots)
    cam = rt.camera(w, h, spp, depth)
    code = {"linear": api.ACCEL_LINEAR, "lbvh": api.ACCEL_LBVH, "grid": api.ACCEL_GRID, "auto": api.ACCEL_AUTO}[accel]
    img = renderer.render(cam, api.make_opts(accel=code))
    ref, seg = O.render(O.scene(scene_id), O.camera(w, h, spp, depth))
    st = renderer.stats()
    assert st.paths == w * h * spp
    assert st.segments == seg
    assert st.accel_used == (code if accel != "auto" else api.ACCEL_GRID)      # the reference's scenes are planar fields
    mism = np.argwhere(bits(img) != bits(ref))
    assert len(mism) == 0, f"{len(mism)} differing channels, first {mism[:3]}"


@pytest.mark.parametrize("w,h", [(64, 40), (33, 17), (5, 1), (7, 3)])
def test_flat_and_by_pixel_finalize_kernels_write_the_same_frame(renderer, w, h):
    """finalize_flat_kernel (rows stay where they are: pairs of values, coalesced) against finalize_kernel (four pixels per
    thread; still used for row placement and for frames that are not 8-byte aligned), even and odd pixel counts, and both
    against the oracle's frame."""
    import torch
    renderer.upload_scene(rt.scene(1))
    cam = rt.camera(w, h, 5, 8)
    ref, _ = O.render(O.scene(1), O.camera(w, h, 5, 8))
    flat = torch.full((h, w, 3), -1.0, dtype=torch.float32, device="cuda:0")
    renderer.render(cam, out=flat)
    backing = torch.full((h * w * 3 + 1,), -1.0, dtype=torch.float32, device="cuda:0")
    odd = backing[1:].view(h, w, 3)                                  # 4 bytes off a 16-byte boundary: the by-pixel kernel
    assert odd.data_ptr() % 8 == 4
    renderer.render(cam, out=odd)
    os.environ["RT_FINALIZE_BY_PIXEL"] = "1"
    try:
        forced = torch.full((h, w, 3), -1.0, dtype=torch.float32, device="cuda:0")
        renderer.render(cam, out=forced)
    finally:
        del os.environ["RT_FINALIZE_BY_PIXEL"]
    torch.cuda.synchronize()
    assert float(backing[0]) == -1.0
    for img in (flat, odd, forced):
        assert np.array_equal(bits(img.cpu().numpy()), bits(ref))


@pytest.mark.parametrize("accel", ["linear", "grid"])
@pytest.mark.parametrize("scene_id,w,h,spp,depth", [(1, 32, 20, 6, 25), (2, 48, 32, 9, 50), (3, 40, 24, 8, 50), (1, 21, 13, 40, 50)])
def test_render_bit_exact_vs_oracle_double(renderer, scene_id, w, h, spp, depth, accel):
    """GlobalDouble path (GD camera.h:133-177): whole frames bit for bit against the oracle, scenes 1-3, through the
    shared-memory scan and through the uniform grid (float walk on the rounded ray, exact tests in double)."""
    renderer.upload_scene(rt.scene(scene_id, double=True))
    cam = rt.camera(w, h, spp, depth, double=True)
    img = renderer.render(cam, api.make_opts(accel=api.ACCEL_GRID if accel == "grid" else api.ACCEL_LINEAR))
    ref, seg = O.render(O.scene(scene_id, True), O.camera(w, h, spp, depth, double=True))
    assert renderer.stats().segments == seg
    assert np.array_equal(bits(img), bits(ref))
    ids, t = renderer.primary_hits(rt.camera(97, 61, double=True), accel=api.ACCEL_GRID if accel == "grid" else api.ACCEL_LINEAR)
    oids, ot = O.primary(O.scene(scene_id, True), O.camera(97, 61, double=True))
    assert np.array_equal(ids, oids) and np.array_equal(bits(t), bits(ot))


def test_double_grid_equals_scan_on_a_frame(renderer):
    """1920x1080, scene 1 in double: grid == scan, bit for bit (2 M pixels, 4 spp)."""
    import torch
    renderer.upload_scene(rt.scene(1, double=True))
    cam = rt.camera(1920, 1080, 4, 50, double=True)
    a = torch.empty((1080, 1920, 3), dtype=torch.float64, device="cuda:0")
    b = torch.empty_like(a)
    renderer.render(cam, api.make_opts(accel=api.ACCEL_LINEAR), out=a)
    seg = renderer.stats().segments
    renderer.render(cam, api.make_opts(accel=api.ACCEL_GRID), out=b)
    assert renderer.stats().segments == seg
    assert torch.equal(a.view(torch.int64), b.view(torch.int64))


def test_render_is_deterministic_and_seeded(renderer):
    renderer.upload_scene(rt.scene(1))
    cam = rt.camera(160, 96, 16, 25)
    a = renderer.render(cam)
    b = renderer.render(cam)
    assert np.array_equal(bits(a), bits(b))
    c = renderer.render(cam, api.make_opts(seed=99))
    assert not np.array_equal(bits(a), bits(c))


@pytest.mark.parametrize("world", [2, 3, 8])
def test_row_split_equals_whole_frame(renderer, world):
    """Interleaved row tiles (RT_SPLIT_ROWS) reassemble to the 1-GPU image bit for bit."""
    renderer.upload_scene(rt.scene(1))
    cam = rt.camera(96, 70, 8, 25)          # 70 rows: last tile is partial
    whole = renderer.render(cam)
    out = np.zeros_like(whole)
    seen = np.zeros(cam.height, dtype=int)
    for rank in range(world):
        o = api.make_opts(split=api.SPLIT_ROWS, rank=rank, world=world, tile_rows=8)
        part = renderer.render(cam, o)
        rows = rt.partition_rows(cam.height, 8, rank, world)
        assert part.shape[0] == len(rows)
        out[rows] = part
        seen[rows] += 1
    assert (seen == 1).all()
    assert np.array_equal(bits(out), bits(whole))


@pytest.mark.parametrize("world", [2, 3, 8])
def test_spp_split_equals_whole_frame(renderer, world):
    """Each rank's int64 accumulation buffer (RT_SPLIT_SPP: its share of the samples) summed in ANY order gives the 1-GPU
    image bit for bit: torch's integer sum (what the NCCL reduce does) and rt_finalize_sum over the list of buffers (what
    the CLI does over NVLink P2P) -- and the buffers are the oracle's accumulators."""
    import torch
    renderer.upload_scene(rt.scene(1))
    w, h, spp = 64, 40, 40
    cam = rt.camera(w, h, spp, 25)
    whole = renderer.render(cam)
    accs = []
    for rank in range(world):
        acc = torch.full((h, w, 3), -7, dtype=torch.int64, device="cuda:0")      # rt_render_partials overwrites
        renderer.render_partials(cam, api.make_opts(split=api.SPLIT_SPP, rank=rank, world=world), acc)
        accs.append(acc)
        s0, s1 = rt.partition_samples(spp, rank, world)
        assert renderer.stats().paths == w * h * (s1 - s0)
    torch.cuda.synchronize()
    want = O.accumulate(O.scene(1), O.camera(w, h, spp, 25), *rt.partition_samples(spp, 1, world))
    assert np.array_equal(accs[1].cpu().numpy(), want)
    total = accs[0].clone()
    for a in reversed(accs[1:]):
        total += a
    assert np.array_equal(bits(renderer.finalize(cam, total)), bits(whole))
    assert np.array_equal(bits(renderer.finalize(cam, accs)), bits(whole))
    ref, _ = O.render(O.scene(1), O.camera(w, h, spp, 25))
    assert np.array_equal(bits(whole), bits(ref))


@pytest.mark.parametrize("knobs", [{"RT_CHUNKS": "1"}, {"RT_CHUNKS": "7"}, {"RT_CHUNKS": "3", "RT_BAND_ROWS": "2"}, {"RT_BAND_ROWS": "1"},
                                   {"RT_BAND_ROWS": "5", "RT_CHUNKS": "40"}, {"RT_NO_TILE_ORDER": "1"}, {"RT_PB_COHORT": "16"}])
def test_image_does_not_depend_on_the_job_partition(renderer, knobs, monkeypatch):
    """Integer accumulation: however the scheduler cuts pixels into bands and samples into jobs (tuning knobs of
    plan_jobs), the frame is the same bit for bit."""
    renderer.upload_scene(rt.scene(3))
    cam = rt.camera(72, 36, 40, 12)                  # 72 x 36: pixels are numbered in 8 x 4 tiles by default
    base = renderer.render(cam)
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    img = renderer.render(cam)
    assert renderer.stats().paths == 72 * 36 * 40
    assert np.array_equal(bits(img), bits(base))
    ref, _ = O.render(O.scene(3), O.camera(72, 36, 40, 12))
    assert np.array_equal(bits(base), bits(ref))


@pytest.mark.parametrize("scene_id", [1, 2, 3])
def test_converged_radiance_vs_reference(renderer, golden_dir, scene_id):
    """MAE <= 1/255 per channel and PSNR >= 40 dB against the reference global-float PPM
    (rendered on a B200 by the reference binary rebuilt for sm_100), 320x192, 4096 spp, 50 bounces."""
    path = os.path.join(golden_dir, f"ref_scene{scene_id}_f32_320x192_4096spp_50b.npz")
    if not os.path.exists(path):
        pytest.skip("reference golden not generated yet")
    ref = np.load(path)["img"]
    renderer.upload_scene(rt.scene(scene_id))
    cam = rt.camera(320, 192, 4096, 50)
    a = rt.ppm_quantise(renderer.render(cam))
    b = rt.ppm_quantise(renderer.render(cam, api.make_opts(seed=4242)))
    mae, psnr = image_metrics(a, ref)
    floor_mae, floor_psnr = image_metrics(a, b)       # new-vs-new two-seed noise floor
    print(f"scene {scene_id}: vs reference MAE {mae} PSNR {psnr:.2f} dB; two-seed floor MAE {floor_mae} PSNR {floor_psnr:.2f} dB")
    assert (mae <= 1.0).all(), (mae, floor_mae)
    assert psnr >= 40.0, (psnr, floor_psnr)
    # a bias would show as an error well above the noise floor
    assert (mae <= floor_mae * 1.5 + 0.1).all(), (mae, floor_mae)


@pytest.mark.parametrize("scene_id", [1, 2, 3])
def test_converged_radiance_vs_reference_double(renderer, golden_dir, scene_id):
    """The double path (rt_render64) against the reference's GlobalDouble PPM (GD camera.h:133-177, material.h:68,
    vec3.h:120-122; rendered on a B200 by the reference binary rebuilt for sm_100): 320x192, 50 bounces, at 4096 spp
    (MAE <= 1/255, PSNR >= 40 dB, as for float) or, where only the round-1 golden exists, at 1024 spp -- there the MAE
    gate of 1/255 sits inside the Monte-Carlo noise of two independent renders (0.8-0.9 code values
    at 1000 spp), so the gate is widened to 1.25 and tied to the measured two-seed noise floor."""
    spp = 4096
    path = os.path.join(golden_dir, f"ref_scene{scene_id}_f64_320x192_{spp}spp_50b.npz")
    if not os.path.exists(path):
        spp = 1024
        path = os.path.join(golden_dir, f"ref_scene{scene_id}_f64_320x192_{spp}spp_50b.npz")
    if not os.path.exists(path):
        pytest.skip("reference golden not generated yet")
    ref = np.load(path)["img"]
    renderer.upload_scene(rt.scene(scene_id, double=True))
    cam = rt.camera(320, 192, spp, 50, double=True)
    a = rt.ppm_quantise(renderer.render(cam).astype(np.float32))
    b = rt.ppm_quantise(renderer.render(cam, api.make_opts(seed=4242)).astype(np.float32))
    mae, psnr = image_metrics(a, ref)
    floor_mae, floor_psnr = image_metrics(a, b)
    print(f"scene {scene_id} double: vs GlobalDouble MAE {mae} PSNR {psnr:.2f} dB; two-seed floor MAE {floor_mae} PSNR {floor_psnr:.2f} dB")
    assert psnr >= 40.0, (psnr, floor_psnr)
    assert (mae <= (1.0 if spp >= 4096 else 1.25)).all(), (mae, floor_mae)   # 1/255; 1024 spp: plus the noise margin
    assert (mae <= floor_mae * 1.25 + 0.1).all(), (mae, floor_mae)   # a bias would sit well above the floor
    assert psnr >= floor_psnr - 1.5, (psnr, floor_psnr)


def test_full_size_properties(renderer):
    """BASELINE config 2 size (1920x1080), reduced spp: path count, finite non-negative output,
    top rows are sky, image responds to the seed only in the noise."""
    renderer.upload_scene(rt.scene(1))
    cam = rt.camera(1920, 1080, 8, 25)
    img = renderer.render(cam)
    st = renderer.stats()
    assert st.paths == 1920 * 1080 * 8
    assert 1.5 < st.segments / st.paths < 4.0
    assert np.isfinite(img).all() and (img >= 0).all() and (img <= 1.0001).all()
    assert img[:40].mean() > 0.6                     # sky gradient
    ids, t = renderer.primary_hits(rt.camera(1920, 1080))
    assert (ids[0] == -1).all() and (ids[-1] >= 0).all()     # top row sky, bottom row ground or a sphere
    assert (ids[-1] == 0).any()


def test_zero_bounces_is_black(renderer):
    renderer.upload_scene(rt.scene(2))
    img = renderer.render(rt.camera(32, 16, 4, 0))
    assert (img == 0).all()


def concentric_scene(n_shells=64):
    """n_shells nested spheres around the look-at point: rays through the middle have far more
    discriminant-positive slots than the per-thread candidate list holds (24), which exercises
    the in-order rescan path; plus the reference's ground sphere and a metal and a glass shell."""
    s = np.zeros(n_shells + 1, dtype=api.SLOT_DTYPE)
    s["c"][0] = (0, -1000, 0)
    s["r"][0] = 1000
    s["albedo"][0] = 0.5
    for k in range(n_shells):
        s["c"][k + 1] = (0, 1, 0)
        s["r"][k + 1] = 0.25 + 0.02 * k
        s["type"][k + 1] = k % 3
        s["albedo"][k + 1] = (0.9, 0.5 + 0.005 * k, 0.3)
        s["fuzz"][k + 1] = 0.1 if k % 3 == 1 else 0
        s["ri"][k + 1] = 1.5 if k % 3 == 2 else 0
    return s


def test_candidate_overflow_rescan_is_exact(renderer):
    slots = concentric_scene()
    renderer.upload_scene(slots)
    cam = rt.camera(80, 48, 6, 12)
    ids, t = renderer.primary_hits(cam)
    oids, ot = O.primary(slots, O.camera(80, 48))
    assert np.array_equal(ids, oids) and np.array_equal(bits(t), bits(ot))
    assert (ids == len(slots) - 1).any()                 # the outermost shell is what a ray meets first
    img = renderer.render(cam)
    ref, seg = O.render(slots, O.camera(80, 48, 6, 12))
    assert renderer.stats().segments == seg
    assert np.array_equal(bits(img), bits(ref))


@pytest.mark.parametrize("n", [1, 31, 32, 33, 64, 65])
def test_slot_counts_around_block_boundaries(renderer, n):
    """The scan walks blocks of 32 slots with a tail mask: every count around the boundaries."""
    slots = rt.scene(1)[:n].copy()
    renderer.upload_scene(slots)
    cam = rt.camera(64, 40, 4, 8)
    ids, t = renderer.primary_hits(cam)
    oids, ot = O.primary(slots, O.camera(64, 40))
    assert np.array_equal(ids, oids) and np.array_equal(bits(t), bits(ot))
    img = renderer.render(cam)
    ref, _ = O.render(slots, O.camera(64, 40, 4, 8))
    assert np.array_equal(bits(img), bits(ref))


# ---------------------------------------------- conservative pre-filter of the paired scan ------
def shifted_scene(scale, shift):
    """Scene 1 scaled and moved away from the coordinate origin: the filter's error bound grows with the
    distance of centres and ray origins from the origin, so this is where a missed candidate would show."""
    s = rt.scene(1).copy()
    s["c"] = (s["c"].astype(np.float64) * scale + np.asarray(shift, dtype=np.float64)).astype(np.float32)
    s["r"] = (s["r"].astype(np.float64) * scale).astype(np.float32)
    return s


AUDIT_SCENES = {
    "scene1": lambda: rt.scene(1), "scene2": lambda: rt.scene(2), "scene3": lambda: rt.scene(3),
    "shifted": lambda: shifted_scene(1.0, (37.0, 3.0, -21.0)),
    "tiny": lambda: shifted_scene(1e-3, (0.004, 0.0, 0.002)),
    "huge": lambda: shifted_scene(4096.0, (0.0, 0.0, 0.0)),
    "concentric": lambda: concentric_scene(),
}


@pytest.mark.parametrize("name", sorted(AUDIT_SCENES))
def test_filter_never_rejects_a_slot_the_reference_accepts(renderer, name):
    """rt_filter_audit: about 10^9 (ray, slot) pairs per scene, a quarter of the rays grazing a silhouette
    within a few ulp.  Not one pair with reference discriminant >= 0 may fail the filter."""
    slots = AUDIT_SCENES[name]()
    renderer.upload_scene(slots)
    a = renderer.filter_audit(rt.camera(320, 192), n_rays=max(1 << 18, (1 << 30) // len(slots)), seed=7)
    assert a["pairs"] > 0, "the filter is expected to be active for this scene"
    assert a["missed"] == 0, a
    assert a["exact_pass"] > 1000 and a["filter_pass"] >= a["exact_pass"]
    # on the reference's scenes the filter must also stay selective: at most 1.5x the reference's own pass count
    # (the camera 13 units away from a millimetre-sized copy of the scene is correct but not selective)
    if name.startswith("scene"):
        assert a["filter_pass"] <= 1.5 * a["exact_pass"], a


@pytest.mark.parametrize("name", ["shifted", "tiny", "huge"])
def test_shifted_and_scaled_scenes_bit_exact_vs_oracle(renderer, name):
    slots = AUDIT_SCENES[name]()
    renderer.upload_scene(slots)
    cam = rt.camera(96, 64, 4, 10)
    ids, t = renderer.primary_hits(cam)
    oids, ot = O.primary(slots, O.camera(96, 64))
    assert np.array_equal(ids, oids) and np.array_equal(bits(t), bits(ot))
    img = renderer.render(cam)
    ref, seg = O.render(slots, O.camera(96, 64, 4, 10))
    assert renderer.stats().segments == seg
    assert np.array_equal(bits(img), bits(ref))


def test_equal_t_ties_go_to_the_lowest_slot(renderer):
    """Duplicates of one sphere in both halves of the slot list and in the far set: the reference's strict
    `t < closest` keeps the first (lowest) slot, whatever order the candidates are resolved in."""
    base = rt.scene(1)
    n = 64
    s = base[:n].copy()
    for dst in (5, 40, 63):                      # half 0, half 1, last slot
        s[dst] = s[2]
    s[50] = s[0]                                 # a second copy of the ground sphere (far set)
    renderer.upload_scene(s)
    cam = rt.camera(160, 96, 4, 8)
    ids, t = renderer.primary_hits(cam)
    oids, ot = O.primary(s, O.camera(160, 96))
    assert np.array_equal(ids, oids) and np.array_equal(bits(t), bits(ot))
    assert (ids == 2).any() and not np.isin(ids, (5, 40, 63, 50)).any()
    img = renderer.render(cam)
    ref, _ = O.render(s, O.camera(160, 96, 4, 8))
    assert np.array_equal(bits(img), bits(ref))


def test_multi_chunk_scene_bit_exact_vs_oracle(renderer):
    """1 448 slots: each half of the slot list spans several chunks of candidate-mask words."""
    slots = rt.scene_scaled(19)
    assert len(slots) > 1024
    renderer.upload_scene(slots)
    cam = rt.camera(120, 72, 3, 8)
    ids, t = renderer.primary_hits(cam)
    oids, ot = O.primary(O.scene_scaled(19), O.camera(120, 72))
    assert np.array_equal(ids, oids) and np.array_equal(bits(t), bits(ot))
    img = renderer.render(cam)
    ref, seg = O.render(O.scene_scaled(19), O.camera(120, 72, 3, 8))
    assert renderer.stats().segments == seg
    assert np.array_equal(bits(img), bits(ref))


def test_scene_outside_the_filter_range_takes_the_exact_scan(renderer):
    """Scenes whose scale would over/underflow the filter bound are scanned with the exact discriminant only."""
    slots = shifted_scene(1e-14, (0.0, 0.0, 0.0))
    renderer.upload_scene(slots)
    a = renderer.filter_audit(rt.camera(64, 40), n_rays=1024)
    assert a["pairs"] == 0
    ids, t = renderer.primary_hits(rt.camera(64, 40))
    oids, ot = O.primary(slots, O.camera(64, 40))
    assert np.array_equal(ids, oids) and np.array_equal(bits(t), bits(ot))


# ------------------------------------------------ camera rays through per-tile candidate lists ------
@pytest.mark.parametrize("scene_id,w,h,spp,depth,double", [(1, 320, 192, 8, 25, False), (2, 200, 120, 8, 50, False),
                                                            (3, 97, 61, 12, 50, False), (1, 64, 40, 6, 25, True)])
def test_primary_bins_equal_the_full_scan(renderer, scene_id, w, h, spp, depth, double):
    """rt_opts.primary_bins: a camera ray resolved against its tile's candidate list gets the hit the shared-memory scan
    returns, so the frames are identical bit for bit -- and equal to the oracle's (which has no bins at all)."""
    renderer.upload_scene(rt.scene(scene_id, double=double))
    cam = rt.camera(w, h, spp, depth, double=double)
    on = renderer.render(cam, api.make_opts(accel=api.ACCEL_LINEAR, primary_bins=api.PBINS_ON))
    st_on = renderer.stats()
    off = renderer.render(cam, api.make_opts(accel=api.ACCEL_LINEAR, primary_bins=api.PBINS_OFF))
    st_off = renderer.stats()
    assert st_on.launches == st_off.launches + 1, "the bin kernel did not run"
    assert (st_on.paths, st_on.segments) == (st_off.paths, st_off.segments)
    assert np.array_equal(bits(on), bits(off))
    if w * h * spp <= 200 * 120 * 8:
        ref, seg = O.render(O.scene(scene_id, double), O.camera(w, h, spp, depth, double=double))
        assert st_on.segments == seg
        assert np.array_equal(bits(on), bits(ref))


@pytest.mark.parametrize("name", ["shifted", "tiny", "huge", "concentric"])
def test_primary_bins_on_awkward_scenes(renderer, name):
    """Far from the origin, millimetre-sized (every slot lands in one tile: the list overflows and the tile's camera
    rays take the scan), 4096x, and 64 nested shells (more candidates than a list holds)."""
    slots = AUDIT_SCENES[name]()
    renderer.upload_scene(slots)
    cam = rt.camera(160, 96, 4, 10)
    on = renderer.render(cam, api.make_opts(accel=api.ACCEL_LINEAR, primary_bins=api.PBINS_ON))
    off = renderer.render(cam, api.make_opts(accel=api.ACCEL_LINEAR, primary_bins=api.PBINS_OFF))
    assert np.array_equal(bits(on), bits(off))


def test_primary_bins_row_split_and_partial_tiles(renderer):
    """Tiles are addressed by GLOBAL pixel coordinates: a row-split rank looks up the same lists; 70 rows and 100
    columns leave partial tiles at both edges."""
    renderer.upload_scene(rt.scene(1))
    cam = rt.camera(100, 70, 8, 25)
    whole = renderer.render(cam, api.make_opts(accel=api.ACCEL_LINEAR, primary_bins=api.PBINS_OFF))
    out = np.zeros_like(whole)
    for rank in range(3):
        o = api.make_opts(split=api.SPLIT_ROWS, rank=rank, world=3, tile_rows=2, accel=api.ACCEL_LINEAR, primary_bins=api.PBINS_ON)
        out[rt.partition_rows(cam.height, 2, rank, 3)] = renderer.render(cam, o)
    assert np.array_equal(bits(out), bits(whole))


def test_primary_bins_full_size_frame(renderer):
    """BASELINE config 4's frame (3840x2160, 32 400 tiles, 33 million camera rays at 4 spp): bins on == bins off."""
    renderer.upload_scene(rt.scene(1))
    cam = rt.camera(3840, 2160, 4, 50)
    on = renderer.render(cam, api.make_opts(accel=api.ACCEL_LINEAR, primary_bins=api.PBINS_ON))
    seg_on = renderer.stats().segments
    off = renderer.render(cam, api.make_opts(accel=api.ACCEL_LINEAR, primary_bins=api.PBINS_OFF))
    assert renderer.stats().segments == seg_on
    assert np.array_equal(bits(on), bits(off))


@pytest.mark.parametrize("scene_id,w,h,spp,depth", [(1, 320, 192, 6, 25), (2, 200, 120, 8, 50), (3, 97, 61, 12, 50)])
def test_primary_bins_lbvh_equal_the_full_traversal(renderer, scene_id, w, h, spp, depth):
    """LBVH scenes: the lists come from bin_kernel_bvh (the bundle walks the tree).  Bins on == bins off == the linear scan
    without bins, bit for bit, and the work counters agree."""
    renderer.upload_scene(rt.scene(scene_id))
    cam = rt.camera(w, h, spp, depth)
    plain = renderer.render(cam, api.make_opts(accel=api.ACCEL_LINEAR, primary_bins=api.PBINS_OFF))
    seg = renderer.stats().segments
    off = renderer.render(cam, api.make_opts(accel=api.ACCEL_LBVH, primary_bins=api.PBINS_OFF))
    st_off = renderer.stats()
    on = renderer.render(cam, api.make_opts(accel=api.ACCEL_LBVH, primary_bins=api.PBINS_ON))
    st_on = renderer.stats()
    assert st_on.segments == st_off.segments == seg
    assert st_on.binned_segments > 0 and st_off.binned_segments == 0
    assert st_on.node_visits < st_off.node_visits
    assert np.array_equal(bits(on), bits(off)) and np.array_equal(bits(on), bits(plain))


def test_primary_bins_lbvh_100k_scene(renderer):
    """BASELINE config 5's scene at config 2's frame size: near tiles are binned, tiles towards the horizon overflow and
    keep the traversal; the frame is the one the plain traversal renders."""
    renderer.upload_scene(rt.scene_scaled(158))
    cam = rt.camera(1920, 1080, 1, 50)
    off = renderer.render(cam, api.make_opts(accel=api.ACCEL_LBVH, primary_bins=api.PBINS_OFF))
    seg = renderer.stats().segments
    on = renderer.render(cam, api.make_opts(accel=api.ACCEL_LBVH, primary_bins=api.PBINS_ON))
    st = renderer.stats()
    assert st.segments == seg
    assert 0.2 * 1920 * 1080 < st.binned_segments <= 1920 * 1080
    assert np.array_equal(bits(on), bits(off))


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_primary_bins_random_scenes(renderer, seed):
    """Random sphere soups (sizes over two decades, some around the camera): bins on == bins off for both structures."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(40, 700))
    s = np.zeros(n, dtype=api.SLOT_DTYPE)
    s["c"] = rng.uniform(-12, 14, size=(n, 3)).astype(np.float32)
    s["r"] = (10 ** rng.uniform(-1.5, 0.5, size=n)).astype(np.float32)
    s["type"] = rng.integers(0, 3, size=n)
    s["albedo"] = rng.uniform(0.2, 0.9, size=(n, 3)).astype(np.float32)
    s["fuzz"] = np.where(s["type"] == 1, 0.2, 0).astype(np.float32)
    s["ri"] = np.where(s["type"] == 2, 1.5, 0).astype(np.float32)
    renderer.upload_scene(s)
    cam = rt.camera(160, 96, 4, 8)
    ref = renderer.render(cam, api.make_opts(accel=api.ACCEL_LINEAR, primary_bins=api.PBINS_OFF))
    for accel in (api.ACCEL_LINEAR, api.ACCEL_LBVH):
        img = renderer.render(cam, api.make_opts(accel=accel, primary_bins=api.PBINS_ON))
        assert np.array_equal(bits(img), bits(ref)), accel


# ------------------------------------------------------------------ LBVH (RT_ACCEL_LBVH) ------
@pytest.mark.parametrize("scene_id", [1, 2, 3])
def test_lbvh_primary_equals_linear_scan(renderer, scene_id):
    renderer.upload_scene(rt.scene(scene_id))
    cam = rt.camera(320, 192)
    ids, t = renderer.primary_hits(cam)
    bids, bt = renderer.primary_hits(cam, accel=api.ACCEL_LBVH)
    assert np.array_equal(ids, bids)
    assert np.array_equal(bits(t), bits(bt))


@pytest.mark.parametrize("scene_id,w,h,spp,depth", [(1, 96, 64, 8, 25), (3, 64, 40, 16, 50)])
def test_lbvh_render_equals_linear_scan(renderer, scene_id, w, h, spp, depth):
    """Same hits => same paths => the same image, bit for bit."""
    renderer.upload_scene(rt.scene(scene_id))
    cam = rt.camera(w, h, spp, depth)
    a = renderer.render(cam, api.make_opts(accel=api.ACCEL_LINEAR, primary_bins=api.PBINS_OFF))       # reference: every segment through the scan
    seg = renderer.stats().segments
    b = renderer.render(cam, api.make_opts(accel=api.ACCEL_LBVH))
    st = renderer.stats()
    assert st.segments == seg and st.node_visits > 0 and 0 < st.sphere_tests < seg * 488
    assert np.array_equal(bits(a), bits(b))


def test_lbvh_mid_size_scene_equals_linear_scan(renderer):
    """3 604 slots still fit the shared-memory scan: LBVH and linear scan must agree exactly."""
    slots = rt.scene_scaled(30)
    assert len(slots) == 3604
    renderer.upload_scene(slots)
    cam = rt.camera(160, 96, 4, 10)
    ids, t = renderer.primary_hits(cam)
    bids, bt = renderer.primary_hits(cam, accel=api.ACCEL_LBVH)
    assert np.array_equal(ids, bids) and np.array_equal(bits(t), bits(bt))
    a = renderer.render(cam, api.make_opts(accel=api.ACCEL_LINEAR, primary_bins=api.PBINS_OFF))       # reference: every segment through the scan
    b = renderer.render(cam, api.make_opts(accel=api.ACCEL_LBVH))
    assert np.array_equal(bits(a), bits(b))


@pytest.mark.parametrize("half", [11, 30])
def test_rays_with_non_finite_directions_hit_nothing_in_every_structure(renderer, half):
    """|d|^2 = +inf or NaN: the reference's roots are (h -/+ sqrt(disc)) / a = +-0 or NaN, never > tmin (GF hittable.h:40-66), so the
    scan reports no hit.  bvh_start answers that without walking the tree -- such a ray passes every slab test and used to visit
    every node (config 5: one lane, 199 719 nodes, ~100 ms at the end of the launch).  Columns of the probe camera: finite rays,
    |d| ~ 1e25 (a overflows), infinite components, NaN components."""
    renderer.upload_scene(rt.scene(1) if half == 11 else rt.scene_scaled(half))
    cam = rt.camera(64, 24)
    first = None
    for i, du in enumerate(((cam.du[0], cam.du[1], cam.du[2]), (1e25, 0.0, -1e24), (float("inf"), 0.0, 0.0), (float("nan"), 1.0, 0.0))):
        for k in range(3):
            cam.du[k] = du[k]
        ids, t = renderer.primary_hits(cam)
        for accel in (api.ACCEL_LBVH, api.ACCEL_GRID):
            aids, at = renderer.primary_hits(cam, accel=accel)
            assert np.array_equal(ids, aids) and np.array_equal(bits(t), bits(at)), (i, accel)
        if i == 0:
            first = ids
            assert (ids >= 0).any()
        else:
            if i == 1:
                assert np.array_equal(ids[:, 0], first[:, 0])              # 0 * 1e25 = 0: column 0 is still the finite camera ray
            assert (ids[:, 1:] == -1).all() and np.isinf(t[:, 1:]).all()   # from column 1 on, |d|^2 is inf or NaN in every row


def test_lbvh_100k_scene_primary_vs_oracle(renderer):
    """BASELINE config 5 scene (99 860 slots): LBVH primary (slot id, t) against the oracle's
    linear hit_world, bit for bit."""
    slots = rt.scene_scaled(158)
    assert len(slots) == 99860
    assert slots.tobytes() == O.scene_scaled(158).tobytes()
    renderer.upload_scene(slots)
    cam = rt.camera(128, 72)
    ids, t = renderer.primary_hits(cam, accel=api.ACCEL_LBVH)
    oids, ot = O.primary(slots, O.camera(128, 72))
    assert np.array_equal(ids, oids)
    assert np.array_equal(bits(t), bits(ot))
    with pytest.raises(rt.RtError):
        renderer.primary_hits(cam)                    # too large for the shared-memory scan
    img = renderer.render(rt.camera(128, 72, 2, 8), api.make_opts(accel=api.ACCEL_LBVH))
    ref, seg = O.render(slots, O.camera(128, 72, 2, 8))
    assert renderer.stats().segments == seg
    assert np.array_equal(bits(img), bits(ref))


def test_lbvh_degenerate_scenes(renderer):
    """1 and 2 slots, duplicates (equal Morton codes), and the concentric shells."""
    base = rt.scene(1)
    for slots in (base[:1].copy(), base[:2].copy(), np.concatenate([base[5:6]] * 7 + [base[:1]]), concentric_scene(40)):
        renderer.upload_scene(slots)
        cam = rt.camera(64, 40)
        ids, t = renderer.primary_hits(cam, accel=api.ACCEL_LBVH)
        oids, ot = O.primary(slots, O.camera(64, 40))
        assert np.array_equal(ids, oids) and np.array_equal(bits(t), bits(ot))


# ------------------------------------------------------------------ wavefront variant ---------
@pytest.mark.parametrize("scene_id,w,h,spp,depth", [(1, 160, 96, 16, 25), (2, 96, 64, 24, 50), (3, 64, 40, 9, 4)])
def test_wavefront_equals_megakernel(renderer, scene_id, w, h, spp, depth):
    """The material-sorted wavefront variant walks the same jobs: identical image, segment and
    path counts."""
    renderer.upload_scene(rt.scene(scene_id))
    cam = rt.camera(w, h, spp, depth)
    a = renderer.render(cam, api.make_opts(accel=api.ACCEL_LINEAR, primary_bins=api.PBINS_OFF))       # reference: every segment through the scan
    sa = renderer.stats()
    b = renderer.render(cam, api.make_opts(kernel=api.KERNEL_WAVEFRONT))
    sb = renderer.stats()
    assert (sa.paths, sa.segments) == (sb.paths, sb.segments) == (w * h * spp, sa.segments)
    assert sb.launches > 3
    assert np.array_equal(bits(a), bits(b))


def test_wavefront_row_split(renderer):
    renderer.upload_scene(rt.scene(1))
    cam = rt.camera(64, 50, 8, 10)
    whole = renderer.render(cam)
    out = np.zeros_like(whole)
    for rank in range(2):
        o = api.make_opts(split=api.SPLIT_ROWS, rank=rank, world=2, tile_rows=4, kernel=api.KERNEL_WAVEFRONT)
        out[rt.partition_rows(50, 4, rank, 2)] = renderer.render(cam, o)
    assert np.array_equal(bits(out), bits(whole))


def test_bench_line_schema():
    """bench.py prints one JSON line with the contract's keys (tiny workload)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", "cfg1", "--steps", "2", "--warmup", "3",
                        "--no-cpu-baseline", "--no-ref-gpu"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-500:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert key in line, key
    assert line["metric"] == "Mpath-samples/s" and line["value"] > 0 and line["gpu_launches"] == 6       # 2 steps x (bin_kernel, trace_kernel_pb, finalize_kernel)
    assert line["e2e"]["h2d_bytes_per_step"] == 488 * 40 and line["e2e"]["d2h_bytes_per_step"] == 320 * 192 * 12
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "reference_equivalent_tflops"} <= set(line["roofline"])
    assert 0 < line["roofline"]["frac"] < 1                    # executed FP32 work can never exceed the peak
    assert "workload" in line["config"] and line["kernel"]["accel"] == "grid"      # scene 1 under rt_opts_default
    lin = line["linear_scan"]                                  # the FP32-bound exhibit on the same workload
    assert lin["accel"] == "linear" and 0 < lin["roofline"]["frac"] < 1 and lin["filter_tests_per_segment"] > 100
    assert line["accel_lbvh"]["accel"] == "lbvh" and line["accel_lbvh"]["node_visits_per_segment"] > 1


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_lbvh_random_scenes_equal_linear_scan(renderer, seed):
    """Random sphere soups (overlaps, radii over 2.5 decades, duplicates, a huge ground sphere):
    any single differing hit would change the image, so LBVH and linear renders must be bit-equal,
    and so must the primary (slot id, t) passes."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(200, 2500))
    s = np.zeros(n, dtype=api.SLOT_DTYPE)
    spread = float(rng.choice([3.0, 15.0, 60.0]))
    s["c"] = rng.uniform(-spread, spread, (n, 3)).astype(np.float32)
    s["c"][:, 1] = np.abs(s["c"][:, 1]) * 0.3
    s["r"] = np.exp(rng.uniform(np.log(0.01), np.log(3.0), n)).astype(np.float32)
    s["type"] = rng.integers(0, 3, n)
    s["albedo"] = rng.uniform(0.2, 1.0, (n, 3)).astype(np.float32)
    s["fuzz"] = np.where(s["type"] == 1, rng.uniform(0, 0.5, n), 0).astype(np.float32)
    s["ri"] = np.where(s["type"] == 2, 1.5, 0).astype(np.float32)
    s["c"][0] = (0, -1000, 0); s["r"][0] = 1000; s["type"][0] = 0
    s[n // 2] = s[n // 3]                                       # an exact duplicate: tie on t, lowest slot wins
    renderer.upload_scene(s)
    cam = rt.camera(96, 64, 6, 12)
    ids, t = renderer.primary_hits(cam)
    bids, bt = renderer.primary_hits(cam, accel=api.ACCEL_LBVH)
    assert np.array_equal(ids, bids) and np.array_equal(bits(t), bits(bt))
    a = renderer.render(cam, api.make_opts(accel=api.ACCEL_LINEAR, primary_bins=api.PBINS_OFF))       # reference: every segment through the scan
    seg = renderer.stats().segments
    b = renderer.render(cam, api.make_opts(accel=api.ACCEL_LBVH))
    assert renderer.stats().segments == seg
    assert np.array_equal(bits(a), bits(b))
    c = renderer.render(cam, api.make_opts(accel=api.ACCEL_AUTO))
    assert np.array_equal(bits(a), bits(c))


def test_cli_end_to_end(tmp_path):
    """The drop-in binary: stdout is `render_ms,e2e_ms` in the reference's format, the PPM has the
    reference's naming scheme and its content is the quantised oracle image, byte for byte."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "raytracingincuda_b200", "bin", "b200-raytrace")
    p = subprocess.run([exe, "--scene_id", "2", "--width", "64", "--height", "40", "--samples", "4", "--bounces", "5",
                        "--threads", "16", "--stats"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    assert re.fullmatch(r" {0,14}\d+\.\d{8}, {0,14}\d+\.\d{8}\n", p.stdout), repr(p.stdout)
    assert len(p.stdout) == 15 + 1 + 15 + 1
    import json
    stats = json.loads(p.stderr.strip().splitlines()[-1])                 # --stats: one JSON object on stderr
    assert stats["paths"] == 64 * 40 * 4 and stats["binned_segments"] == stats["paths"] and stats["segments"] > stats["paths"]
    name = "b200_float_scene2_64x40_4samples_5bounces_16threadsPerBlockRow.ppm"
    assert os.listdir(tmp_path) == [name]
    ref, _ = O.render(O.scene(2), O.camera(64, 40, 4, 5))
    want = tmp_path / "want.ppm"
    rt.ppm_write(str(want), ref)
    assert (tmp_path / name).read_bytes() == want.read_bytes()
    # double precision variant and the prefix override
    p = subprocess.run([exe, "--scene_id", "3", "--width", "32", "--height", "20", "--samples", "2", "--bounces", "3",
                        "--precision", "double", "--prefix", "global_double_"], cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode == 0 and (tmp_path / "global_double_scene3_32x20_2samples_3bounces_8threadsPerBlockRow.ppm").exists()
    ref64, _ = O.render(O.scene(3, True), O.camera(32, 20, 2, 3, double=True))
    got = np.array((tmp_path / "global_double_scene3_32x20_2samples_3bounces_8threadsPerBlockRow.ppm").read_text().split()[4:], dtype=int)
    x = np.clip(ref64, 0.0, 0.999)
    assert np.array_equal(got, (256 * x).astype(int).reshape(-1))


def test_cli_multi_gpu_splits_write_the_single_gpu_ppm(tmp_path):
    """`--gpus N --split rows|spp` (P2P and host gather) write the PPM the single-GPU run writes, byte for byte; unknown
    values of the extension options fail like a malformed cxxopts argument (abort, status 134)."""
    import subprocess
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "raytracingincuda_b200", "bin", "b200-raytrace")
    common = ["--scene_id", "1", "--width", "100", "--height", "70", "--samples", "10", "--bounces", "12"]
    p = subprocess.run([exe, *common, "--accel", "bogus"], cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode == -6 or p.returncode == 134, p.returncode            # SIGABRT
    assert "incorrect_argument_type" in p.stderr
    for bad in (["--split", "cols"], ["--primary_bins", "maybe"], ["--precision", "half"], ["--seed", "x1"], ["--kernel", "fast"]):
        assert subprocess.run([exe, *common, *bad], cwd=tmp_path, capture_output=True).returncode in (-6, 134), bad
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    name = "b200_float_scene1_100x70_10samples_12bounces_8threadsPerBlockRow.ppm"
    one = tmp_path / "one"
    one.mkdir()
    assert subprocess.run([exe, *common], cwd=one, capture_output=True).returncode == 0
    want = (one / name).read_bytes()
    n = min(torch.cuda.device_count(), 4)
    for k, extra in enumerate((["--split", "rows"], ["--split", "spp"], ["--split", "rows", "--gather", "host"],
                               ["--split", "spp", "--gather", "host"], ["--split", "spp", "--gather", "p2p-load"],
                               ["--split", "spp", "--accel", "linear"])):
        d = tmp_path / f"multi{k}"
        d.mkdir()
        p = subprocess.run([exe, *common, "--gpus", str(n), *extra, "--stats"], cwd=d, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, (extra, p.stderr[-500:])
        assert (d / name).read_bytes() == want, extra
        assert f'"gpus": {n}' in p.stderr and f'"split": "{extra[1]}"' in p.stderr


def test_benchmark_driver_writes_the_reference_csv_schema(tmp_path):
    """tools/benchmark.py: the reference's sweep (global_float_benchmark.sh:25-82) and its averaging step
    (timing-benchmarks/process.py:16-33) over the drop-in binary -- same CSV header, one row per run, group means."""
    import csv
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "raytracingincuda_b200", "bin", "b200-raytrace")
    out = tmp_path / "b200.csv"
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "benchmark.py"), "--exe", exe, "--out", str(out), "--scenes", "1",
                        "--sizes", "64x40,96x64", "--samples", "4", "--bounces", "5", "--threads", "8,16", "--runs", "2", "--", "--no-ppm"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-500:]
    rows = list(csv.reader(open(out)))
    assert rows[0] == ["scene_id", "width", "height", "samples", "bounces", "threads", "run", "render_only_time_ms", "end_to_end_time_ms"]
    assert len(rows) == 1 + 2 * 2 * 2                             # threads x sizes x runs
    assert all(float(r[7]) > 0 and float(r[8]) >= float(r[7]) for r in rows[1:])
    avg = list(csv.reader(open(tmp_path / "b200_avg.csv")))
    assert avg[0][:6] == ["scene_id", "width", "height", "samples", "bounces", "threads"] and len(avg) == 1 + 4
    k = [r for r in rows[1:] if r[:6] == avg[1][:6]]
    assert abs(float(avg[1][6]) - sum(float(r[7]) for r in k) / len(k)) < 1e-3


def test_cli_scene_file_round_trip(tmp_path):
    """General scene loader: --dump_scene writes the generated slots, --scene_file renders them back to the same PPM;
    a hand-written three-sphere file renders to the oracle's image."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "raytracingincuda_b200", "bin", "b200-raytrace")
    common = ["--width", "64", "--height", "40", "--samples", "4", "--bounces", "6"]
    a, b = tmp_path / "a", tmp_path / "b"
    a.mkdir(); b.mkdir()
    p = subprocess.run([exe, "--scene_id", "3", *common, "--dump_scene", str(tmp_path / "s3.txt")], cwd=a, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    p = subprocess.run([exe, "--scene_id", "3", *common, "--scene_file", str(tmp_path / "s3.txt")], cwd=b, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    name = "b200_float_scene3_64x40_4samples_6bounces_8threadsPerBlockRow.ppm"
    assert (a / name).read_bytes() == (b / name).read_bytes()
    (tmp_path / "mini.txt").write_text("# ground, a glass ball, a metal ball\n0 -1000 0 1000 0 0.5 0.5 0.5 0 0\n"
                                        "0 1 0 1 2 0 0 0 0 1.5\n4 1 0 1 1 0.7 0.6 0.5 0 0\n")
    c = tmp_path / "c"
    c.mkdir()
    p = subprocess.run([exe, "--scene_id", "9", *common, "--scene_file", str(tmp_path / "mini.txt")], cwd=c, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    slots = rt.load_scene(tmp_path / "mini.txt")
    assert len(slots) == 3 and list(slots["type"]) == [0, 2, 1]
    ref, _ = O.render(slots, O.camera(64, 40, 4, 6))
    rt.ppm_write(str(tmp_path / "want.ppm"), ref)
    assert (c / "b200_float_scene9_64x40_4samples_6bounces_8threadsPerBlockRow.ppm").read_bytes() == (tmp_path / "want.ppm").read_bytes()


@pytest.mark.parametrize("w,h", [(96, 70), (50, 33)])
def test_place_rows_writes_the_full_frame_in_place(renderer, w, h):
    """rt_opts.place_rows: each rank stores its rows at their global positions of one device frame
    (what the CLI does across GPUs over NVLink P2P); width % 4 != 0 takes the per-pixel store path."""
    import torch
    renderer.upload_scene(rt.scene(1))
    cam = rt.camera(w, h, 6, 10)
    whole = renderer.render(cam)
    frame = torch.full((h, w, 3), -1.0, dtype=torch.float32, device="cuda:0")
    for rank in range(3):
        o = api.make_opts(split=api.SPLIT_ROWS, rank=rank, world=3, tile_rows=2)
        o.place_rows = 1
        renderer.render(cam, o, out=frame)
    torch.cuda.synchronize()
    assert np.array_equal(bits(frame.cpu().numpy()), bits(whole))
    with pytest.raises(rt.RtError):
        o = api.make_opts(split=api.SPLIT_ROWS, rank=0, world=2)
        o.place_rows = 1
        renderer.render(cam, o, out=np.zeros((h, w, 3), dtype=np.float32))     # host frame: rejected


def test_config4_frame_size_properties(renderer):
    """BASELINE config 4's frame (3840x2160) at reduced spp: path/segment accounting, an 8-way row
    split placed straight into one frame equals the whole-frame render, the LBVH render equals the
    linear-scan render, and the primary pass agrees between both structures -- all bit for bit."""
    import torch
    W, H = 3840, 2160
    renderer.upload_scene(rt.scene(1))
    cam = rt.camera(W, H, 4, 50)
    whole = torch.empty((H, W, 3), dtype=torch.float32, device="cuda:0")
    renderer.render(cam, out=whole)
    st = renderer.stats()
    assert st.paths == W * H * 4 and st.chunks == rt.num_chunks(W, H, 4) and 2.0 < st.segments / st.paths < 4.0
    placed = torch.zeros_like(whole)
    for rank in range(8):
        o = api.make_opts(split=api.SPLIT_ROWS, rank=rank, world=8)
        o.place_rows = 1
        renderer.render(cam, o, out=placed)
    assert torch.equal(placed.view(torch.int32), whole.view(torch.int32))
    lb = torch.empty_like(whole)
    for accel in (api.ACCEL_LINEAR, api.ACCEL_LBVH, api.ACCEL_GRID):
        renderer.render(cam, api.make_opts(accel=accel), out=lb)
        assert renderer.stats().segments == st.segments
        assert torch.equal(lb.view(torch.int32), whole.view(torch.int32)), accel
    # 8-way spp split: eight int64 accumulation buffers, added inside rt_finalize_sum
    cam8 = rt.camera(W, H, 8, 50)
    renderer.render(cam8, out=whole)
    accs = []
    for rank in range(8):
        acc = torch.empty((H, W, 3), dtype=torch.int64, device="cuda:0")
        renderer.render_partials(cam8, api.make_opts(split=api.SPLIT_SPP, rank=rank, world=8), acc)
        accs.append(acc)
    renderer.finalize(cam8, accs, out=lb)
    assert torch.equal(lb.view(torch.int32), whole.view(torch.int32))
    del accs
    renderer.render(cam, out=whole)
    img = whole.cpu().numpy()
    assert np.isfinite(img).all() and img.min() >= 0 and img.max() <= 1.0001
    ids, t = renderer.primary_hits(rt.camera(W, H))
    bids, bt = renderer.primary_hits(rt.camera(W, H), accel=api.ACCEL_LBVH)
    assert np.array_equal(ids, bids) and np.array_equal(bits(t), bits(bt))
    assert len(np.unique(ids)) > 100                      # the 20-degree view sees about 140 of the 488 slots


@pytest.mark.parametrize("w,h,spp,depth,double", [(1920, 1080, 100, 25, False), (3840, 2160, 1000, 50, False),
                                                   (1920, 1080, 100, 50, True)])
def test_full_size_frame_spot_check_vs_oracle(renderer, w, h, spp, depth, double):
    """BASELINE configs 2, 3 and 4 at FULL size (config 4: 8.3 G path-samples, 265 M jobs -- the size where the job decode's
    64-bit multiply-high divisions and the band / tail-region arithmetic matter): the frame is rendered once on the device
    and random pixels, the four corners and the first / last pixel of bands are recomputed sample by sample with the
    oracle (orc_sample + integer accumulation) and compared bit for bit."""
    import torch
    scene_id = 1 if not double else 2
    renderer.upload_scene(rt.scene(scene_id, double=double))
    cam = rt.camera(w, h, spp, depth, double=double)
    frame = torch.empty((h, w, 3), dtype=torch.float64 if double else torch.float32, device="cuda:0")
    renderer.render(cam, out=frame)
    st = renderer.stats()
    assert st.paths == w * h * spp
    rng = np.random.default_rng(w + spp)
    n_rand = 40 if spp <= 100 else 20
    pts = [(0, 0), (w - 1, 0), (0, h - 1), (w - 1, h - 1), (w // 2, h // 2), (w - 1, h // 2), (0, h // 3)]
    pts += [(int(rng.integers(0, w)), int(rng.integers(0, h))) for _ in range(n_rand)]
    oslots, ocam = O.scene(scene_id, double), O.camera(w, h, spp, depth, double=double)
    ys = torch.tensor([p[1] for p in pts], device="cuda:0")
    xs = torch.tensor([p[0] for p in pts], device="cuda:0")
    got = frame[ys, xs].cpu().numpy()
    for k, (i, j) in enumerate(pts):
        want = O.pixel(oslots, ocam, i, j)
        assert np.array_equal(bits(got[k]), bits(want)), (i, j, got[k], want)


CHECKED_CHILD = r"""
import sys, json
import numpy as np
import torch
import raytracingincuda_b200 as rt
from raytracingincuda_b200 import api
r = rt.Renderer(0)
en, code, n = r.debug_checks()
assert en, "not the checked build"
report = {}
def run(label, fn):
    fn()
    en, code, n = r.debug_checks()
    report[label] = [code, n]
L, B, G, A = api.ACCEL_LINEAR, api.ACCEL_LBVH, api.ACCEL_GRID, api.ACCEL_AUTO
for sid in (1, 2, 3):
    r.upload_scene(rt.scene(sid))
    cam = rt.camera(97, 61, 9, 25)
    for accel, name in ((L, "linear"), (B, "lbvh"), (G, "grid")):
        run(f"scene{sid} {name}", lambda: r.render(cam, api.make_opts(accel=accel)))
        run(f"scene{sid} {name} primary", lambda: r.primary_hits(rt.camera(97, 61), accel=accel))
    run(f"scene{sid} linear bins off", lambda: r.render(cam, api.make_opts(accel=L, primary_bins=api.PBINS_OFF)))
    run(f"scene{sid} lbvh bins off", lambda: r.render(cam, api.make_opts(accel=B, primary_bins=api.PBINS_OFF)))
    run(f"scene{sid} wavefront", lambda: r.render(cam, api.make_opts(accel=L, kernel=api.KERNEL_WAVEFRONT)))
    for rank in range(3):
        run(f"scene{sid} rows {rank}/3", lambda: r.render(cam, api.make_opts(split=api.SPLIT_ROWS, rank=rank, world=3, tile_rows=4)))
    acc = torch.empty((61, 97, 3), dtype=torch.int64, device="cuda:0")
    for rank in range(2):
        run(f"scene{sid} spp {rank}/2", lambda: r.render_partials(cam, api.make_opts(split=api.SPLIT_SPP, rank=rank, world=2), acc))
    run(f"scene{sid} finalize", lambda: r.finalize(cam, acc))
    r.upload_scene(rt.scene(sid, double=True))
    run(f"scene{sid} double", lambda: r.render(rt.camera(64, 40, 6, 25, double=True)))
    run(f"scene{sid} double primary", lambda: r.primary_hits(rt.camera(64, 40, double=True)))
for half, label in ((24, "2 308 slots"), (158, "99 860 slots")):
    r.upload_scene(rt.scene_scaled(half))
    cam = rt.camera(160, 96, 3, 12)
    for accel, name in ((B, "lbvh"), (G, "grid"), (A, "auto")) + (((L, "linear"),) if half == 24 else ()):
        run(f"{label} {name}", lambda: r.render(cam, api.make_opts(accel=accel)))
        run(f"{label} {name} primary", lambda: r.primary_hits(rt.camera(160, 96), accel=accel))
r.upload_scene(rt.scene(1))
run("1080p frame, 2 spp", lambda: r.render(rt.camera(1920, 1080, 2, 50), out=torch.empty((1080, 1920, 3), dtype=torch.float32, device="cuda:0")))
selftest = r.debug_checks(selftest=True)
print(json.dumps({"report": report, "selftest": list(selftest)}))
"""


def test_checked_build_finds_no_out_of_bounds_access():
    """compute-sanitizer is closed on this GPU pool (profiles/logs/r02_sanitizer_closed_message.txt), so the memcheck role is
    played by the CHECKED build of the library (-DRT_CHECKS=1: a bounds assertion at every indexed access of the kernels):
    every kernel family on every scene kind must finish without one violation, and the assertion machinery must be alive
    (its self-test violates one assertion on purpose)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "raytracingincuda_b200", "librt_b200_checked.so")
    assert os.path.exists(lib), "build it with make -C raytracingincuda_b200/csrc"
    env = dict(os.environ, RT_B200_LIB=lib, PYTHONPATH=root)
    p = subprocess.run([sys.executable, "-c", CHECKED_CHILD], env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    out = json.loads(p.stdout.strip().splitlines()[-1])
    bad = {k: v for k, v in out["report"].items() if v[1] != 0}
    assert not bad, bad
    assert len(out["report"]) > 60
    assert out["selftest"] == [True, 999, 1]


def test_production_build_has_no_checks(renderer):
    assert renderer.debug_checks() == (False, 0, 0)


def test_invalid_inputs_fail_loudly(renderer):
    """Empty scene, bad material type, zero-sized frame, render before upload, precision mismatch."""
    with pytest.raises(rt.RtError):
        renderer.upload_scene(np.zeros(0, dtype=api.SLOT_DTYPE))
    bad = rt.scene(2).copy()
    bad["type"][3] = 7
    with pytest.raises(rt.RtError):
        renderer.upload_scene(bad)
    fresh = rt.Renderer(0)
    try:
        with pytest.raises(rt.RtError) as e:
            fresh.render(rt.camera(8, 8, 1, 1))
        assert e.value.code == -2                          # RT_ENOSCENE
        fresh.upload_scene(rt.scene(2, double=True))
        with pytest.raises(rt.RtError) as e:
            fresh.render(rt.camera(8, 8, 1, 1))            # float call on a double scene
        assert e.value.code == -6                          # RT_EPRECISION
    finally:
        fresh.close()
    renderer.upload_scene(rt.scene(2))
    cam = rt.camera(8, 8, 1, 1)
    cam.width = 0
    with pytest.raises(rt.RtError):
        renderer.render(cam, out=np.zeros((8, 8, 3), dtype=np.float32))
    one = renderer.render(rt.camera(1, 1, 1, 1))           # the smallest frame there is
    assert one.shape == (1, 1, 3) and np.isfinite(one).all()
    with pytest.raises(rt.RtError) as e:
        renderer.render(rt.camera(8, 8, 1, 1), api.make_opts(primary_bins=7))
    assert e.value.code == -1                              # RT_EINVAL


# ------------------------------------------------------------------ uniform grid (RT_ACCEL_GRID) ------
# The algorithm is also stated operation by operation in float32 in tools/grid_model.py and checked on the CPU against the
# oracle's hit_world (tests/test_grid_model.py).
GRID_SCENES = {"scene1": lambda: rt.scene(1), "scene2": lambda: rt.scene(2), "scene3": lambda: rt.scene(3),
               "shifted": lambda: shifted_scene(1.0, (37.0, 3.0, -21.0)), "scaled24": lambda: rt.scene_scaled(24),
               "scaled158": lambda: rt.scene_scaled(158)}


def test_grid_is_refused_for_a_scene_that_is_not_a_field(renderer):
    """One sphere plus the ground: no two similar spheres -> RT_EINVAL for RT_ACCEL_GRID, while AUTO renders it."""
    s = rt.scene(1)[:2].copy()
    s["r"][1] = 30.0
    renderer.upload_scene(s)
    with pytest.raises(rt.RtError) as e:
        renderer.render(rt.camera(16, 16, 1, 2), api.make_opts(accel=api.ACCEL_GRID))
    assert e.value.code == -1                              # RT_EINVAL
    img = renderer.render(rt.camera(16, 16, 1, 2))
    assert renderer.stats().accel_used == api.ACCEL_LINEAR and np.isfinite(img).all()


def test_auto_picks_grid_lbvh_linear(renderer):
    """RT_ACCEL_AUTO: planar fields of similar spheres -> grid, from the reference's scenes up to the 99 860-slot field of BASELINE
    config 5 (the grid is 6-9 % ahead of the LBVH there; bench.py still reports config 5 through the LBVH the config names);
    a 3-D soup -> LBVH; tiny scenes -> linear scan."""
    cam = rt.camera(32, 20, 1, 4)
    for slots, want in ((rt.scene(1), api.ACCEL_GRID), (rt.scene(3), api.ACCEL_GRID), (rt.scene(2), api.ACCEL_GRID),
                        (rt.scene(1)[:20].copy(), api.ACCEL_LINEAR), (rt.scene_scaled(158), api.ACCEL_GRID),
                        (rt.scene_scaled(12), api.ACCEL_GRID), (rt.scene_scaled(60), api.ACCEL_GRID)):
        renderer.upload_scene(slots)
        renderer.render(cam)
        assert renderer.stats().accel_used == want, (len(slots), renderer.stats().accel_used)
    rng = np.random.default_rng(5)
    s = np.zeros(600, dtype=api.SLOT_DTYPE)
    s["c"] = rng.uniform(-8, 8, (600, 3)).astype(np.float32)
    s["r"] = 0.2
    s["albedo"] = 0.5
    renderer.upload_scene(s)
    renderer.render(cam)
    assert renderer.stats().accel_used == api.ACCEL_LBVH
    renderer.upload_scene(rt.scene(1, double=True))
    renderer.render(rt.camera(32, 20, 1, 4, double=True))
    assert renderer.stats().accel_used == api.ACCEL_GRID            # double: float walk on the rounded ray, exact tests in double
    with pytest.raises(rt.RtError):
        renderer.render(rt.camera(32, 20, 1, 4, double=True), api.make_opts(accel=api.ACCEL_LBVH))   # the LBVH is a float structure


@pytest.mark.parametrize("name", sorted(GRID_SCENES))
def test_grid_primary_equals_linear_scan(renderer, name):
    slots = GRID_SCENES[name]()
    renderer.upload_scene(slots)
    cam = rt.camera(320, 192)
    gids, gt = renderer.primary_hits(cam, accel=api.ACCEL_GRID)
    ids, t = renderer.primary_hits(cam, accel=api.ACCEL_LBVH if len(slots) > 60000 else api.ACCEL_LINEAR)
    assert np.array_equal(gids, ids) and np.array_equal(bits(gt), bits(t))


@pytest.mark.parametrize("scene_id,w,h,spp,depth", [(1, 320, 192, 8, 25), (2, 200, 120, 8, 50), (3, 97, 61, 12, 50)])
def test_grid_render_equals_linear_scan(renderer, scene_id, w, h, spp, depth):
    renderer.upload_scene(rt.scene(scene_id))
    cam = rt.camera(w, h, spp, depth)
    ref = renderer.render(cam, api.make_opts(accel=api.ACCEL_LINEAR, primary_bins=api.PBINS_OFF))
    seg = renderer.stats().segments
    img = renderer.render(cam, api.make_opts(accel=api.ACCEL_GRID))
    st = renderer.stats()
    assert st.segments == seg and 0 < st.sphere_tests < seg * 40
    assert np.array_equal(bits(img), bits(ref))


@pytest.mark.parametrize("seed", [21, 22, 23, 24, 25, 26])
def test_grid_random_fields_equal_linear_scan(renderer, seed):
    """Random fields of similar spheres (jittered lattice, radii within a factor 3, tilted or raised slabs, a few odd-sized
    spheres, an exact duplicate, the ground): grid and linear renders and primary passes must be bit-equal; spheres of
    negative radius (the hollow-glass idiom: r*r in the reference's arithmetic) included."""
    rng = np.random.default_rng(seed)
    side = int(rng.integers(6, 40))
    n = side * side + 4
    s = np.zeros(n, dtype=api.SLOT_DTYPE)
    gx, gz = np.meshgrid(np.arange(side), np.arange(side))
    pitch = float(rng.choice([0.6, 1.0, 2.5]))
    s["c"][:side * side, 0] = (gx.ravel() - side / 2) * pitch + rng.uniform(-0.4, 0.4, side * side) * pitch
    s["c"][:side * side, 2] = (gz.ravel() - side / 2) * pitch + rng.uniform(-0.4, 0.4, side * side) * pitch
    s["c"][:side * side, 1] = 0.2 + rng.uniform(0, float(rng.choice([0.0, 0.5, 3.0])), side * side)
    s["r"][:side * side] = 0.2 * np.exp(rng.uniform(np.log(0.6), np.log(1.8), side * side))
    s["r"][3] = -s["r"][3]                                      # negative radius
    s["c"][-4] = (0, -1000, 0); s["r"][-4] = 1000
    s["c"][-3] = (0, 1, 0); s["r"][-3] = 1.0
    s["c"][-2] = (-4, 1, 0); s["r"][-2] = 0.004                  # far smaller than the field's spheres
    s[-1] = s[7]                                                 # exact duplicate: tie on t, lowest slot wins
    s["type"] = rng.integers(0, 3, n)
    s["type"][-4] = 0
    s["albedo"] = rng.uniform(0.2, 1.0, (n, 3)).astype(np.float32)
    s["fuzz"] = np.where(s["type"] == 1, rng.uniform(0, 0.5, n), 0).astype(np.float32)
    s["ri"] = np.where(s["type"] == 2, 1.5, 0).astype(np.float32)
    renderer.upload_scene(s)
    cam = rt.camera(96, 64, 6, 12)
    ids, t = renderer.primary_hits(cam)
    for accel in (api.ACCEL_GRID, api.ACCEL_LBVH):
        gids, gt = renderer.primary_hits(cam, accel=accel)
        assert np.array_equal(ids, gids) and np.array_equal(bits(t), bits(gt)), accel
    a = renderer.render(cam, api.make_opts(accel=api.ACCEL_LINEAR, primary_bins=api.PBINS_OFF))
    seg = renderer.stats().segments
    for accel in (api.ACCEL_GRID, api.ACCEL_LBVH, api.ACCEL_AUTO):
        b = renderer.render(cam, api.make_opts(accel=accel))
        assert renderer.stats().segments == seg
        assert np.array_equal(bits(a), bits(b)), accel
