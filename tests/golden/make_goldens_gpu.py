"""Generate the golden fixtures that need a B200: outputs of the REFERENCE itself.

Run on the GPU box (the reference-derived binaries under oracle/_ref/ travel with the snapshot;
/root/reference itself does not exist there):
    python tests/golden/make_goldens_gpu.py gpurun_out/golden [--only-missing]
then copy gpurun_out/golden/*.npz into tests/golden/ and commit.

  primary_scene{1,2,3}_{f32,f64}.npz   (slot id, t) of the reference's own hit_world() for the
                                       deterministic primary pass (oracle/ref_harness.cu)
  ref_scene{1,2,3}_{f32,f64}_WxH_Sspp_Bb.npz   8-bit code values of the PPM written by the
                                       reference binary rebuilt for sm_100 (+ its render_ms)
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")


def read_ppm(path):
    with open(path, "rb") as f:
        tok = f.read().split()
    assert tok[0] == b"P3"
    w, h, mx = int(tok[1]), int(tok[2]), int(tok[3])
    return np.array(tok[4:], dtype=np.int32).reshape(h, w, 3).astype(np.uint8)


def main(out_dir, only_missing=False):
    os.makedirs(out_dir, exist_ok=True)
    tmp = tempfile.mkdtemp()
    for sid in (() if only_missing else (1, 2, 3)):
        for tag, suffix, W, H in (("f32", "float", 320, 192), ("f64", "double", 160, 96)):
            dump = os.path.join(tmp, f"scene{sid}_{tag}.dump")
            subprocess.check_call([os.path.join(REF, f"scene_dump_{suffix}"), "--scene_id", str(sid)],
                                  env=dict(os.environ, ORC_SCENE_DUMP=dump))
            raw = os.path.join(tmp, "primary.bin")
            subprocess.check_call([os.path.join(REF, f"ref_harness_{suffix}"), "primary", str(W), str(H), dump, raw])
            n = W * H
            ids = np.fromfile(raw, dtype=np.int32, count=n)
            t = np.fromfile(raw, dtype=np.float32 if tag == "f32" else np.float64, offset=4 * n, count=n)
            np.savez_compressed(os.path.join(out_dir, f"primary_scene{sid}_{tag}.npz"),
                                ids=ids.reshape(H, W).astype(np.int16), t=t.reshape(H, W))
            print("primary", sid, tag, "hits", int((ids >= 0).sum()), "of", n, flush=True)
    jobs = [(1, "float", 320, 192, 4096, 50), (2, "float", 320, 192, 4096, 50), (3, "float", 320, 192, 4096, 50),
            (1, "double", 320, 192, 1024, 50), (1, "float", 320, 192, 10, 25), (1, "float", 320, 192, 100, 25),
            # round 2: the GlobalDouble renders of all three scenes at the spp the float goldens use
            (1, "double", 320, 192, 4096, 50), (2, "double", 320, 192, 4096, 50), (3, "double", 320, 192, 4096, 50)]
    for sid, prec, W, H, spp, b in jobs:
        name = f"ref_scene{sid}_{'f32' if prec == 'float' else 'f64'}_{W}x{H}_{spp}spp_{b}b.npz"
        if only_missing and os.path.exists(os.path.join(HERE, name)):
            continue
        exe = os.path.join(REF, f"global-{prec}-cuda-raytrace")
        out = subprocess.check_output([exe, "--scene_id", str(sid), "--width", str(W), "--height", str(H),
                                       "--samples", str(spp), "--bounces", str(b), "--threads", "8"], cwd=tmp)
        render_ms, e2e_ms = [float(x) for x in out.decode().strip().split(",")]
        ppm = os.path.join(tmp, f"global_{prec}_scene{sid}_{W}x{H}_{spp}samples_{b}bounces_8threadsPerBlockRow.ppm")
        img = read_ppm(ppm)
        tag = "f32" if prec == "float" else "f64"
        np.savez_compressed(os.path.join(out_dir, f"ref_scene{sid}_{tag}_{W}x{H}_{spp}spp_{b}b.npz"),
                            img=img, render_ms=render_ms, e2e_ms=e2e_ms)
        print("ref render", sid, prec, W, H, spp, b, "render_ms", render_ms, "Msamp/s",
              W * H * spp / render_ms / 1e3, flush=True)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if a != "--only-missing"]
    main(args[0] if args else "gpurun_out/golden", only_missing="--only-missing" in sys.argv)
