"""Pin a whole bench frame on the UNMODIFIED reference (oracle/_ref) -- TEST INFRASTRUCTURE, offline.

    python tools/make_headline_pin.py [--workload balls] [--threads 7] [--rows all|N]

Renders the frame `bench.py --workload NAME` times (same scene, camera, lights, pixelfactor, max_lvl) with
the reference's own raytracing.cpp/mesh.cpp behind oracle/ref_harness.cpp, row set by row set (checkpointed
under gpurun_out/pin_<name>/, so an interrupted run resumes), and writes tests/golden/pins/<name>.npz:

  rows            int32[R]      the pixel rows that were rendered (all of them for the headline frame)
  id_crc          uint32[R]     zlib.crc32 of the row's per-sample primary primitive ids (int32, W*pfx*pfy of them,
                                sample order ((y*W+x)*pfx+subx)*pfy+suby -- the order of main.cpp:369-390)
  ids_z           bytes         the same ids, all rendered rows, int32 little endian, zlib level 9 (so a failing
                                test can say WHICH samples differ, not only which rows)
  u8              uint8[R,W,3]  the quantised rows (main.cpp:116-117 truncation of the clamped average)
  u8_crc          uint32[R]     crc32 of each u8 row
  rgb_crc         uint32[R]     crc32 of the float32 bits of each clamped RGB row
  + W, H, pf, max_lvl, eye, corners, lights, n_triangles, scene_crc

About 1.5 h of 8 cores for the 800x800x16 headline frame (24 M rays x 44,672 triangles + the id pass).
"""
import argparse
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402
from raytracert_b200 import host  # noqa: E402
import bench  # noqa: E402


def scene_crc(scene):
    c = 0
    for a in (scene.vertices, scene.indices, scene.tri_material, scene.materials):
        c = zlib.crc32(np.ascontiguousarray(a).tobytes(), c)
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="balls")
    ap.add_argument("--threads", type=int, default=max(1, (os.cpu_count() or 2) - 1))
    ap.add_argument("--rows", default="all", help="'all' or a row count: that many rows spread evenly over the frame")
    ap.add_argument("--group", type=int, default=16, help="rows per checkpointed call")
    args = ap.parse_args()
    scene, W, H, pf, lvl, eye, center, lights, desc = bench.workload(args.workload)
    cam = host.Camera(W, H, eye, center)
    lights = np.asarray([cam.eye] if lights is None else lights, np.float32)
    rows = np.arange(H) if args.rows == "all" else np.unique(np.linspace(0, H - 1, int(args.rows)).round().astype(int))
    R = pyoracle.RefOracle()
    R.set_scene(scene)
    R.configure(cam.eye, lights, 63, lvl)
    ck = os.path.join(ROOT, "gpurun_out", f"pin_{args.workload}")
    os.makedirs(ck, exist_ok=True)
    spp = pf * pf
    t0 = time.time()
    # the harness renders lattices y0::ystep; a row set {y} is rendered as y0 = y, ystep = H (one row per call) in
    # groups so a checkpoint holds `group` rows
    for g0 in range(0, len(rows), args.group):
        grp = rows[g0:g0 + args.group]
        path = os.path.join(ck, f"rows_{grp[0]:05d}_{grp[-1]:05d}.npz")
        if os.path.exists(path):
            continue
        ids = np.zeros((len(grp), W * spp), np.int32)
        rgb = np.zeros((len(grp), W, 3), np.float32)
        for i, y in enumerate(grp):
            f_rgb, _, f_prim = R.render(cam.corners, W, H, pf, pf, y0=int(y), ystep=H, want_samples=True, threads=args.threads)
            ids[i] = f_prim.reshape(H, W * spp)[y]
            rgb[i] = f_rgb[y]
        np.savez(path + ".tmp.npz", rows=grp, ids=ids, rgb=rgb)
        os.replace(path + ".tmp.npz", path)
        done = g0 + len(grp)
        el = time.time() - t0
        print(f"{done}/{len(rows)} rows, {el:.0f} s", flush=True)
    all_ids, all_rgb = [], []
    for g0 in range(0, len(rows), args.group):
        grp = rows[g0:g0 + args.group]
        z = np.load(os.path.join(ck, f"rows_{grp[0]:05d}_{grp[-1]:05d}.npz"))
        assert np.array_equal(z["rows"], grp)
        all_ids.append(z["ids"]); all_rgb.append(z["rgb"])
    ids = np.concatenate(all_ids); rgb = np.concatenate(all_rgb)
    u8 = R.quantise(rgb)
    out = os.path.join(ROOT, "tests", "golden", "pins")
    os.makedirs(out, exist_ok=True)
    np.savez_compressed(
        os.path.join(out, args.workload + ".npz"), rows=rows.astype(np.int32),
        id_crc=np.array([zlib.crc32(r.astype("<i4").tobytes()) for r in ids], np.uint32),
        ids_z=np.frombuffer(zlib.compress(ids.astype("<i4").tobytes(), 9), np.uint8),
        u8=u8, u8_crc=np.array([zlib.crc32(r.tobytes()) for r in u8], np.uint32),
        rgb_crc=np.array([zlib.crc32(r.astype("<f4").tobytes()) for r in rgb], np.uint32),
        W=W, H=H, pf=pf, max_lvl=lvl, eye=cam.eye, corners=cam.corners, lights=lights, n_triangles=scene.n_triangles,
        scene_crc=np.uint32(scene_crc(scene)), kind=np.array("reference (oracle/_ref: unmodified raytracing.cpp + mesh.cpp)"))
    print("wrote", os.path.join(out, args.workload + ".npz"), "hit fraction", float(np.mean(ids >= 0)))


if __name__ == "__main__":
    main()
