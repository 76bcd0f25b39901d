"""CPU-only sweep of the round-2 filters at FULL scene size (no GPU): reflection (mirror) pencils and thread pencils, replayed by
tests/pencil_check.cpp (the same rt_pencil.h / rt_tpencil.h code the CUDA library compiles, fmaf() for FFMA) against the
oracle's decision for EVERY (continuation ray, triangle) pair of

  * the headline scene (Balls stand-in, 44,672 triangles) under the bench camera and a low camera: level-1 continuation rays of a
    lattice of the frame's primary rays, built exactly like reflection() / addOffset() build them (float32);
  * the 1 M-triangle sphere of BASELINE configs[3] with a ground (a smaller lattice).

Prints one JSON line (committed as profiles/r2_reflection_soundness_sweep.json).  Minutes of CPU time.

  python tools/reflection_soundness_sweep.py [--rays 20000] [--sphere-rays 600]
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C
import numpy as np
import test_pencil_filter as T
from oracle import pyoracle
from raytracert_b200 import host, scenes

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=20000)
ap.add_argument("--sphere-rays", type=int, default=600)
args = ap.parse_args()
port = pyoracle.PortOracle()
if not os.path.exists(T.SO):
    raise SystemExit("run `python -m pytest tests/test_pencil_filter.py -k thread` once to build the replay library")
L = C.CDLL(T.SO)
L.pencil_check.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.POINTER(T.Result)]
L.pencil_check_set_plane.argtypes = [C.c_double] * 4
L.pencil_check_set_premise.argtypes = [C.c_int]
L.tpencil_check.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(T.Result)]
pair_fn = C.cast(port.L.orc_ray_triangle, C.c_void_p)


def sweep(name, scene, cam, W, H, pf, n_rays, plane=None):
    t0 = time.time()
    tris = np.ascontiguousarray(T.tri_array(scene), np.float32).reshape(-1, 9)
    M = T.magnitude_bound(scene, cam.corners)
    step = max(1, int(np.sqrt(W * H * pf * pf / max(n_rays, 1))))
    # blocks of neighbouring samples (several rays per facet, like the sub-samples of a pixel): a lattice of 3x3 sample blocks
    rays = T.primary_rays(cam.corners, W, H, pf, 1) if W * H * pf * pf <= 4 * n_rays else None
    if rays is None:
        full = T.primary_rays(cam.corners, W, H, pf, step)
        rays = full
    port.set_scene(scene)
    port.configure(cam.eye, np.zeros((0, 3), np.float32), 0, 0)
    _, prim, hit = port.trace(rays[:, :3], rays[:, 3:])
    ok = prim >= 0
    brays = np.ascontiguousarray(T.reflected_rays(rays[ok], hit[ok], scene.normals[prim[ok]]), np.float32)
    refl = np.ascontiguousarray(prim[ok], np.int32)
    out = {"scene": name, "triangles": int(len(tris)), "continuation_rays": int(len(brays)), "launches": []}
    t = tris.reshape(-1, 3)
    lo = (t.min(axis=0) - 0.01).astype(np.float32); hi = (t.max(axis=0) + 0.01).astype(np.float32)
    eye = np.ascontiguousarray(cam.eye, np.float64)
    r = T.Result()
    L.tpencil_check(eye.ctypes.data, 6e-6, float(M), lo.ctypes.data, hi.ctypes.data, len(tris), tris.ctypes.data, len(brays), brays.ctypes.data, refl.ctypes.data, 8,
                    pair_fn, C.byref(r))
    acc = max(1, len(brays) - r.unsafe_rays)
    out["launches"].append({"kind": "thread pencils (rays grouped by reflector, 8 per thread)", "setup_ok": bool(r.setup_ok), "rays": int(len(brays)),
                            "refused_by_the_acceptance_check": int(r.unsafe_rays), "pairs": int(r.pairs), "accepted_by_reference": int(r.ref_hits),
                            "hot_candidates": int(r.candidates), "full_test_survivors": int(r.grazing_skipped), "violations": int(r.violations), "delta": r.delta,
                            "hot_candidates_per_ray": r.candidates / acc, "full_test_survivors_per_ray": r.grazing_skipped / acc})
    if plane is not None:
        n, d = plane
        tt = tris.reshape(-1, 3, 3)[refl]
        on = np.all(np.abs(tt @ np.asarray(n, np.float64) - d) < 1e-6, axis=1)
        L.pencil_check_set_plane(float(n[0]), float(n[1]), float(n[2]), float(d))
        L.pencil_check_set_premise(1)
        corners = np.ascontiguousarray(cam.corners, np.float32)
        for label, batch in (("mirror pencil: rays reflected off the plane group", brays[on]), ("mirror pencil: rays reflected off other surfaces (must be refused or harmless)", brays[~on])):
            batch = np.ascontiguousarray(batch, np.float32)
            if len(batch) == 0:
                continue
            r = T.Result()
            L.pencil_check(3, corners.ctypes.data, float(M), len(tris), tris.ctypes.data, len(batch), batch.ctypes.data, 1.0, pair_fn, C.byref(r))
            out["launches"].append({"kind": label, "plane": [float(x) for x in n] + [float(d)], "setup_ok": bool(r.setup_ok), "rays": int(len(batch)),
                                    "refused_by_the_acceptance_check": int(r.unsafe_rays), "pairs": int(r.pairs), "accepted_by_reference": int(r.ref_hits),
                                    "candidates": int(r.candidates), "violations": int(r.violations), "delta": r.delta, "cos_g": r.cos_g})
    out["seconds"] = time.time() - t0
    out["pairs_total"] = sum(l["pairs"] for l in out["launches"])
    out["violations_total"] = sum(l["violations"] for l in out["launches"])
    return out


res = []
s = scenes.balls_standin()
res.append(sweep("Balls stand-in (headline scene), bench camera", s, host.Camera(800, 800, (0.0, 2.6, 5.2), (0.0, 0.55, 0.0)), 800, 800, 4, args.rays, plane=((0.0, 1.0, 0.0), 0.0)))
res.append(sweep("Balls stand-in, low camera (near-tangent rays over the terrain)", s, host.Camera(800, 500, (0.2, 0.75, 4.6), (0.0, 0.62, 0.0)), 800, 500, 4, args.rays // 2,
                 plane=((0.0, 1.0, 0.0), 0.0)))
if args.sphere_rays > 0:
    s1 = scenes.tessellated_sphere(ground=True)
    res.append(sweep("1 M-triangle sphere (configs[3]) + ground", s1, host.Camera(3840, 2160, (0.0, 0.6, 3.4), (0, 0, 0)), 3840, 2160, 4, args.sphere_rays))
print(json.dumps({"what": "CPU replay of the reflection (mirror) pencils and the thread pencils against the oracle, every (continuation ray, triangle) pair, full scene size",
                  "sweeps": res, "pairs_total": sum(r["pairs_total"] for r in res), "violations_total": sum(r["violations_total"] for r in res)}))
