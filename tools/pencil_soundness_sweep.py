"""CPU-only sweep of the pencil filter's soundness at full scene size (no GPU): replays the filter (tests/pencil_check.cpp,
the same record code the CUDA library compiles) against the oracle's decision for EVERY (ray, triangle) pair of

  * the headline scene (Balls stand-in, 44,672 triangles): a lattice of the frame's primary rays, rays aimed at edges and
    vertices, shadow rays from the oracle's hit points and from edge-aligned origins, for the bench light and two more;
  * the 1 M-triangle sphere of BASELINE configs[3] (a smaller ray lattice).

Prints one JSON line (committed as profiles/r1h_pencil_soundness_sweep.json).  Minutes of CPU time.

  python tools/pencil_soundness_sweep.py [--rays 6000] [--sphere-rays 400]
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
ap.add_argument("--rays", type=int, default=6000)
ap.add_argument("--sphere-rays", type=int, default=400)
args = ap.parse_args()

port = pyoracle.PortOracle()
if not os.path.exists(T.SO):
    raise SystemExit("run `python -m pytest tests/test_pencil_filter.py` once to build the replay library")
L = C.CDLL(T.SO)
L.pencil_check.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.POINTER(T.Result)]
pair_fn = C.cast(port.L.orc_ray_triangle, C.c_void_p)


def run(mode, setup, M, tris, rays, scale=1.0):
    setup = np.ascontiguousarray(setup, np.float32); tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
    r = T.Result()
    L.pencil_check(mode, setup.ctypes.data, float(M), len(tris), tris.ctypes.data, len(rays), rays.ctypes.data, scale, pair_fn, C.byref(r))
    return r


def sweep(name, scene, cam, W, H, pf, lights, n_rays):
    t0 = time.time()
    tris = T.tri_array(scene)
    M = T.magnitude_bound(scene, cam.corners)
    step = max(1, int(np.sqrt(W * H * pf * pf / max(n_rays, 1))))
    rays = T.primary_rays(cam.corners, W, H, pf, step)
    rng = np.random.default_rng(3)
    P = T.edge_points(tris, rng, max(200, n_rays // 4))
    eye = np.asarray(cam.eye, np.float32)
    d = P - eye
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    fwd = np.asarray(cam.corners, np.float32).reshape(4, 2, 3)
    fwd = (fwd[:, 1] - fwd[:, 0]).mean(axis=0)
    keep = d @ fwd > 0
    adv = np.concatenate([eye + d[keep], eye + np.float32(9.0) * d[keep]], axis=1).astype(np.float32)
    out = {"scene": name, "triangles": int(len(tris)), "launches": []}
    for label, batch in (("primary rays of the frame", rays), ("primary rays aimed at edges / vertices", adv)):
        for scale in (1.0, 1.0 + 3 * 2.0 ** -24):
            r = run(0, cam.corners, M, tris, batch, scale)
            out["launches"].append({"kind": label, "chart_scale": scale, "setup_ok": bool(r.setup_ok), "rays": int(len(batch)), "pairs": int(r.pairs),
                                    "accepted_by_reference": int(r.ref_hits), "candidates": int(r.candidates), "violations": int(r.violations),
                                    "grazing_skipped": int(r.grazing_skipped), "delta": r.delta, "cos_g": r.cos_g})
    port.set_scene(scene)
    port.configure(cam.eye, np.zeros((0, 3), np.float32), 0, 0)
    _, prim, hit = port.trace(rays[:, :3], rays[:, 3:])
    origins = (hit[prim >= 0] + np.float32(0.1)).astype(np.float32)
    lo, hi = T.scene_box(tris, M)
    for Lp in lights:
        Lp = np.asarray(Lp, np.float32)
        Pe = T.edge_points(tris, rng, max(200, n_rays // 4))
        back = (Pe + (Pe - Lp) * rng.uniform(0.05, 1.5, (len(Pe), 1)).astype(np.float32)).astype(np.float32)
        for label, o in (("shadow rays from the oracle's hit points", origins), ("shadow rays through edges / vertices", back)):
            srays = np.concatenate([o, np.broadcast_to(Lp, o.shape)], axis=1)
            r = run(1, np.concatenate([Lp, lo, hi]).astype(np.float32), M, tris, srays)
            out["launches"].append({"kind": label, "light": [float(x) for x in Lp], "setup_ok": bool(r.setup_ok), "rays": int(len(srays)),
                                    "outside_chart": int(r.unsafe_rays), "pairs": int(r.pairs), "accepted_by_reference": int(r.ref_hits),
                                    "candidates": int(r.candidates), "violations": int(r.violations), "grazing_skipped": int(r.grazing_skipped), "cos_g": r.cos_g})
    out["seconds"] = time.time() - t0
    out["pairs_total"] = sum(l["pairs"] for l in out["launches"])
    out["violations_total"] = sum(l["violations"] for l in out["launches"])
    return out


L.generic_check.argtypes = [C.c_int, C.c_double, C.c_float, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.POINTER(T.Result)]


def generic_sweep(name, scene, cam, W, H, pf, lights, n_rays, clause_free):
    """The generic filter (pencil_check.cpp:generic_check) on frame rays, continuation-ray shaped rays and shadow rays."""
    t0 = time.time()
    tris = np.ascontiguousarray(T.tri_array(scene), np.float32)
    M = T.magnitude_bound(scene, cam.corners)
    rng = np.random.default_rng(23)
    step = max(1, int(np.sqrt(W * H * pf * pf / max(n_rays, 1))))
    rays = T.primary_rays(cam.corners, W, H, pf, step)
    port.set_scene(scene)
    port.configure(cam.eye, np.zeros((0, 3), np.float32), 0, 0)
    _, prim, hit = port.trace(rays[:, :3], rays[:, 3:])
    origins = (hit[prim >= 0] + np.float32(0.1)).astype(np.float32)
    batches = [("primary rays of the frame", rays), ("continuation-ray shaped rays, half of them aimed at edges / vertices", T.bounce_like_rays(tris, rng, n_rays))]
    for Lp in lights:
        Lp = np.asarray(Lp, np.float32)
        Pe = T.edge_points(tris, rng, max(200, n_rays // 4))
        back = (Pe + (Pe - Lp) * rng.uniform(0.05, 1.5, (len(Pe), 1)).astype(np.float32)).astype(np.float32)
        o = np.concatenate([origins, back])
        batches.append((f"shadow rays to {tuple(float(x) for x in Lp)}", np.concatenate([o, np.broadcast_to(Lp, o.shape)], axis=1)))
    out = {"scene": name + " -- GENERIC filter", "triangles": int(len(tris)), "launches": []}
    for label, batch in batches:
        batch = np.ascontiguousarray(batch, np.float32)
        bmin = -0.5 if clause_free and T.grazing_product(tris, batch) * 1.11e-5 < 1e-5 else 1.0e-5
        for mode in (0, 1):
            r = T.Result()
            L.generic_check(mode, float(M), bmin, len(tris), tris.ctypes.data, len(batch), batch.ctypes.data, 1.0 + 2.0 ** -22, 1.0 - 2.0 ** -22, pair_fn, C.byref(r))
            out["launches"].append({"kind": label, "mode": "nearest-hit (tightest distance bound)" if mode == 0 else "any-hit", "grazing_clause": bmin > 0, "setup_ok": True,
                                    "rays": int(len(batch)), "pairs": int(r.pairs), "accepted_by_reference": int(r.ref_hits), "candidates": int(r.candidates),
                                    "violations": int(r.violations), "grazing_skipped": 0})
    out["seconds"] = time.time() - t0
    out["pairs_total"] = sum(l["pairs"] for l in out["launches"])
    out["violations_total"] = sum(l["violations"] for l in out["launches"])
    return out


res = []
s = scenes.balls_standin()
cam = host.Camera(800, 800, (0.0, 2.6, 5.2), (0.0, 0.55, 0.0))
res.append(sweep("Balls stand-in, bench camera", s, cam, 800, 800, 4, [(2.5, 4.0, 3.0), (0.0, 0.0, 4.0), (-6.0, 3.0, 1.0)], args.rays))
res.append(generic_sweep("Balls stand-in, bench camera", s, cam, 800, 800, 4, [(2.5, 4.0, 3.0)], args.rays // 2, True))
cam2 = host.Camera(800, 800)
res.append(sweep("Balls stand-in, default camera (eye in the water plane)", s, cam2, 800, 800, 4, [tuple(cam2.eye)], args.rays // 2))
if args.sphere_rays > 0:
    s1 = scenes.tessellated_sphere()
    cam3 = host.Camera(3840, 2160, (0.0, 0.6, 3.4), (0, 0, 0))
    res.append(sweep("1 M-triangle sphere (configs[3])", s1, cam3, 3840, 2160, 4, [(2.5, 4.0, 3.0)], args.sphere_rays))
print(json.dumps({"what": "CPU replay of the pencil filter (and restatement of the generic filter) against the oracle, every (ray, triangle) pair", "sweeps": res,
                  "pairs_total": sum(r["pairs_total"] for r in res), "violations_total": sum(r["violations_total"] for r in res)}))
