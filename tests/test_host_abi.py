"""Host half of the product (no GPU needed): the C-ABI library loads and exports every symbol
include/rt_b200.h declares; the scene generator, camera and PPM writer are bit-exact against the
reference goldens; the CLI reproduces the reference's flag handling."""
import ctypes as C
import json
import os
import re
import subprocess

import numpy as np
import pytest

import raytracingincuda_b200 as rt
from raytracingincuda_b200 import api
import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "raytracingincuda_b200", "bin", "b200-raytrace")
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "global-float-cuda-raytrace")


def header_functions():
    src = open(os.path.join(ROOT, "include", "rt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 25
    L = C.CDLL(api.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in rt_b200.h but not exported"
    assert sorted(api.SYMBOLS) == names, "api.SYMBOLS and the header drifted apart"
    assert rt.lib().rt_abi_version() == 3


def test_struct_layouts_match_header():
    assert C.sizeof(api.Camera) == 96 and C.sizeof(api.Camera64) == 176
    assert api.SLOT_DTYPE.itemsize == 40 and api.SLOT64_DTYPE.itemsize == 80
    assert C.sizeof(api.Opts) == 64


@pytest.mark.parametrize("scene_id", [1, 2, 3])
@pytest.mark.parametrize("double", [False, True])
def test_scene_generator_bit_exact(golden_dir, scene_id, double):
    s = rt.scene(scene_id, double)
    g = np.fromfile(os.path.join(golden_dir, f"scene{scene_id}{'_f64' if double else ''}.bin"), dtype=s.dtype)
    assert s.tobytes() == g.tobytes()


def test_scene_generator_does_not_touch_process_rand():
    libc = C.CDLL("libc.so.6")
    libc.srand(12345)
    a = libc.rand()
    libc.srand(12345)
    rt.scene(1)
    assert libc.rand() == a


def test_scene_capacity_and_default_id():
    buf = np.zeros(10, dtype=api.SLOT_DTYPE)
    n = rt.lib().rt_scene_generate(1, buf.ctypes.data, 10)
    assert n == 488 and buf.tobytes() == rt.scene(1)[:10].tobytes()
    assert rt.scene(42).tobytes() == rt.scene(3).tobytes()
    big = rt.scene_scaled(158)
    assert len(big) == 99860 and big[0]["r"] == 1000 and (big["r"] > 0).sum() > 99000
    assert rt.scene_scaled(11).tobytes() == rt.scene(1).tobytes()


def test_camera_bit_exact(golden_dir):
    cams = json.load(open(os.path.join(golden_dir, "camera.json")))
    for key, ref in cams.items():
        wh, prec = key.split("_")
        w, h = map(int, wh.split("x"))
        cam = rt.camera(w, h, double=(prec == "f64"))
        for f in ("center", "pixel00", "du", "dv", "disk_u", "disk_v"):
            assert list(getattr(cam, f)) == ref[f], (key, f)
        assert cam.defocus_angle == ref["defocus_angle"]
    assert rt.camera(320, 192, 10, 25).scale == np.float32(1.0) / np.float32(10)
    with pytest.raises(rt.RtError):
        rt.camera(0, 10)


def test_partitions():
    for h, tile, world in [(2160, 8, 8), (70, 8, 3), (5, 8, 4), (1080, 4, 2)]:
        seen = np.concatenate([rt.partition_rows(h, tile, r, world) for r in range(world)])
        assert sorted(seen) == list(range(h))
        for r in range(world):
            rows = rt.partition_rows(h, tile, r, world)
            assert (np.diff(rows) > 0).all()
            assert all((j // tile) % world == r for j in rows)
    for spp, world in [(8, 1), (8, 2), (1000, 4), (1000, 8), (1000, 3), (72, 8), (5, 8)]:
        b = [rt.partition_samples(spp, r, world) for r in range(world)]
        assert b[0][0] == 0 and b[-1][1] == spp and all(x[1] == y[0] for x, y in zip(b, b[1:]))
        assert max(x[1] - x[0] for x in b) - min(x[1] - x[0] for x in b) <= 1
    with pytest.raises(rt.RtError):
        rt.partition_rows(10, 8, 2, 2)
    with pytest.raises(rt.RtError):
        rt.partition_samples(10, 2, 2)
    assert rt.num_chunks(3840, 2160, 1000) == 1000 and rt.num_chunks(320, 192, 4096) == 4096 and rt.num_chunks(8, 8, 10 ** 6) == 65536
    for w, h, spp in [(3840, 2160, 1000), (1920, 1080, 100), (320, 192, 10), (320, 192, 4096), (320, 192, 5), (8, 8, 100000),
                      (7680, 4320, 100000), (640, 360, 17), (3840, 2160, 256), (97, 61, 12)]:
        assert rt.num_chunks(w, h, spp) == O.num_chunks(w, h, spp), (w, h, spp)


def test_ppm_writer_matches_reference_format(tmp_path, golden_dir):
    rng = np.random.default_rng(0)
    img = rng.uniform(-0.1, 1.2, size=(7, 5, 3)).astype(np.float32)
    img[0, 0] = (0.0, 0.999, 1.0)
    p = tmp_path / "x.ppm"
    rt.ppm_write(str(p), img)
    want = "P3\n5 7\n255\n" + "".join(
        f"{int(np.float32(256) * min(max(r, np.float32(0)), np.float32(0.999)))} "
        f"{int(np.float32(256) * min(max(g, np.float32(0)), np.float32(0.999)))} "
        f"{int(np.float32(256) * min(max(b, np.float32(0)), np.float32(0.999)))}\n"
        for r, g, b in img.reshape(-1, 3))
    assert p.read_text() == want
    assert np.array_equal(rt.ppm_quantise(img).reshape(-1), np.array(want.split()[4:], dtype=np.uint8))
    with pytest.raises(rt.RtError):
        rt.ppm_write(str(tmp_path / "no" / "such" / "dir.ppm"), img)


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the device half fails loudly (RT_ENODEVICE), it never renders."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(rt.RtError) as e:
        rt.Renderer(0)
    assert e.value.code == -5


def run(cmd, cwd):
    return subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)


def test_cli_help_and_errors(tmp_path):
    r = run([CLI, "--help"], tmp_path)
    assert r.returncode == 0 and r.stdout.startswith("Super Raytrace: Raytracing with CUDA\nUsage:\n  ./cuda-raytrace [OPTION...]")
    assert "--threads arg   Number of threads per 2-D thread block row. (default: \n                      8)" in r.stdout
    r = run([CLI], tmp_path)
    assert r.returncode == 1 and r.stderr == "Error: --scene_id is required.\n" and "Usage:" in r.stdout
    r = run([CLI, "--bogus"], tmp_path)
    assert r.returncode == -6 and "Option 'bogus' does not exist" in r.stderr      # SIGABRT, shell rc 134
    r = run([CLI, "--scene_id", "1", "--width", "abc"], tmp_path)
    assert r.returncode == -6


@pytest.mark.skipif(not os.path.exists(REF_CLI), reason="reference binary not built (no /root/reference)")
def test_cli_matches_reference_binary_text(tmp_path):
    for argv in (["--help"], [], ["-h"]):
        mine, ref = run([CLI] + argv, tmp_path), run([REF_CLI] + argv, tmp_path)
        assert (mine.returncode, mine.stdout, mine.stderr) == (ref.returncode, ref.stdout, ref.stderr)
    mine, ref = run([CLI, "--nope", "1"], tmp_path), run([REF_CLI, "--nope", "1"], tmp_path)
    assert mine.returncode == ref.returncode and mine.stderr == ref.stderr


def test_ppm_compare_tool(tmp_path):
    """The scalar comparator of the parity gate: MAE / PSNR / max-abs over code values, diff image."""
    exe = os.path.join(ROOT, "raytracingincuda_b200", "bin", "ppm-compare")
    rng = np.random.default_rng(3)
    a = rng.uniform(0, 1, (9, 11, 3)).astype(np.float32)
    b = np.clip(a + rng.normal(0, 0.003, a.shape), 0, 1).astype(np.float32)
    c = np.clip(a + 0.2, 0, 1).astype(np.float32)
    for name, img in (("a", a), ("b", b), ("c", c)):
        rt.ppm_write(str(tmp_path / f"{name}.ppm"), img)
    r = run([exe, "a.ppm", "b.ppm", "d.ppm"], tmp_path)
    out = json.loads(r.stdout)
    qa, qb = rt.ppm_quantise(a).astype(int), rt.ppm_quantise(b).astype(int)
    assert r.returncode == 0 and out["within_tolerance"]
    assert np.allclose(out["mae"], np.abs(qa - qb).mean(axis=(0, 1)), atol=1e-5)
    assert out["max_abs"] == np.abs(qa - qb).max()
    mse = ((qa - qb) ** 2).mean()
    assert abs(out["psnr_db"] - 10 * np.log10(255 ** 2 / mse)) < 1e-3
    diff = np.array((tmp_path / "d.ppm").read_text().split()[4:], dtype=int).reshape(9, 11, 3)
    assert np.array_equal(diff, np.abs(qa - qb))
    r = run([exe, "a.ppm", "c.ppm"], tmp_path)
    assert r.returncode == 3 and not json.loads(r.stdout)["within_tolerance"]
    assert run([exe, "a.ppm", "missing.ppm"], tmp_path).returncode == 1


def test_scene_text_round_trip(tmp_path):
    """rt_scene_write_text / rt_scene_read_text (general scene loader): %.9g round-trips float32, so all three
    reference scenes come back byte for byte; comments and blank lines are skipped."""
    for sid in (1, 2, 3):
        slots = rt.scene(sid)
        path = tmp_path / f"scene{sid}.txt"
        rt.save_scene(path, slots)
        with open(path, "a") as f:
            f.write("\n   # trailing comment\n")
        back = rt.load_scene(path)
        assert back.tobytes() == slots.tobytes()


def test_scene_text_errors(tmp_path):
    with pytest.raises(rt.RtError) as e:
        rt.load_scene(tmp_path / "missing.txt")
    assert e.value.code == -4                                   # RT_EIO
    bad = tmp_path / "bad.txt"
    bad.write_text("0 0 0 1 0 0.5 0.5 0.5 0 0\n1 2 3\n")
    with pytest.raises(rt.RtError) as e:
        rt.load_scene(bad)
    assert e.value.code == -1                                   # RT_EINVAL: short line
    bad.write_text("0 0 0 1 7 0.5 0.5 0.5 0 0\n")
    with pytest.raises(rt.RtError) as e:
        rt.load_scene(bad)
    assert e.value.code == -1                                   # RT_EINVAL: unknown material type
