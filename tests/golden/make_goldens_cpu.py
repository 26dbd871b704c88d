"""Generate the golden fixtures that can be produced WITHOUT a GPU, from the reference itself.

Run in the build container (needs /root/reference and `make -C oracle ref`):
    python tests/golden/make_goldens_cpu.py

  scene{1,2,3}.bin / scene{1,2,3}_f64.bin
      canonical slot records extracted from the host->device payloads
      of the UNMODIFIED reference main.cu (oracle/_ref/scene_dump_*; oracle/ref_scene_hook.h).
      Fields the reference leaves uninitialised for a material type are zeroed.
  camera.json
      camera::initialize() of the reference (oracle/_ref/ref_harness_* camera W H).
  cpu_320x192_10spp_25b.md5
      md5 of the serial CPU reference's PPM at BASELINE config 1.
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import SLOT_DTYPE, SLOT64_DTYPE  # noqa: E402


def read_dump(path):
    blobs = []
    with open(path, "rb") as f:
        while True:
            h = f.read(8)
            if len(h) < 8:
                break
            n = int(np.frombuffer(h, dtype="<u8")[0])
            blobs.append(f.read(n))
    return blobs


def canonical(dump_path, double):
    mats_b, sph_b = read_dump(dump_path)[:2]
    if double:
        mdt = np.dtype([("type", "<i4"), ("pad", "<i4"), ("albedo", "<f8", 3), ("fuzz", "<f8"), ("ri", "<f8")])
        sdt = np.dtype([("c", "<f8", 3), ("r", "<f8"), ("mat", "<u8")])
        out_dt = SLOT64_DTYPE
    else:
        mdt = np.dtype([("type", "<i4"), ("albedo", "<f4", 3), ("fuzz", "<f4"), ("ri", "<f4")])
        sdt = np.dtype([("c", "<f4", 3), ("r", "<f4"), ("mat", "<u8")])
        out_dt = SLOT_DTYPE
    mats = np.frombuffer(mats_b, dtype=mdt)
    sph = np.frombuffer(sph_b, dtype=sdt)
    assert len(mats) == len(sph)
    out = np.zeros(len(sph), dtype=out_dt)
    never = sph["mat"] == 0          # a written slot always carries &h_materials[i]
    out["c"] = sph["c"]
    out["r"] = np.where(never, 0, sph["r"])
    t = np.where(never, 0, mats["type"])
    out["type"] = t
    out["albedo"] = np.where((t < 2)[:, None] & ~never[:, None], mats["albedo"], 0)
    out["fuzz"] = np.where((t == 1) & ~never, mats["fuzz"], 0)
    out["ri"] = np.where((t == 2) & ~never, mats["ri"], 0)
    return out, int(never.sum()), sph, mats


def main():
    meta = {}
    for sid in (1, 2, 3):
        for double in (False, True):
            tag = "d" if double else "f"
            dump = f"/tmp/scene{sid}_{tag}.dump"
            exe = os.path.join(REF, "scene_dump_double" if double else "scene_dump_float")
            subprocess.check_call([exe, "--scene_id", str(sid)], env=dict(os.environ, ORC_SCENE_DUMP=dump))
            rec, n_never, sph, mats = canonical(dump, double)
            name = f"scene{sid}_f64.bin" if double else f"scene{sid}.bin"
            rec.tofile(os.path.join(HERE, name))
            raw_never = [(int(i), sph[i].tobytes().hex(), mats[i].tobytes().hex())
                         for i in np.nonzero(sph["mat"] == 0)[0]]
            meta[name] = {"slots": len(rec), "never_written": n_never,
                          "sha256": hashlib.sha256(rec.tobytes()).hexdigest(),
                          "never_written_raw": raw_never}
            print(name, meta[name]["slots"], n_never, meta[name]["sha256"])
    cams = {}
    for (w, h) in ((320, 192), (160, 96), (1920, 1080), (3840, 2160), (64, 40)):
        for double in (False, True):
            exe = os.path.join(REF, "ref_harness_double" if double else "ref_harness_float")
            cams[f"{w}x{h}_{'f64' if double else 'f32'}"] = json.loads(
                subprocess.check_output([exe, "camera", str(w), str(h)]))
    with open(os.path.join(HERE, "camera.json"), "w") as f:
        json.dump(cams, f, indent=1)
    with open(os.path.join(HERE, "scenes.json"), "w") as f:
        json.dump(meta, f, indent=1)
    ppm = subprocess.run([os.path.join(REF, "inoneweekend_cpu"), "320", "192", "10", "25"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    md5 = hashlib.md5(ppm.stdout).hexdigest()
    with open(os.path.join(HERE, "cpu_320x192_10spp_25b.md5"), "w") as f:
        f.write(md5 + "\n")
    print("cpu baseline md5", md5, "ms", ppm.stderr.decode().strip())


if __name__ == "__main__":
    main()
