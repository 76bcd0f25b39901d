"""Soundness of the pencil filter (raytracert_b200/csrc/rt_pencil.h) on the CPU.

The pencil filter of k_trace / k_shadow uses only IEEE FMAs and adds, so tests/pencil_check.cpp replays it with fmaf()
-- the same record construction code the CUDA library compiles, the same operation order -- for EVERY (ray, triangle)
pair of a scene and compares with the oracle's decision for that pair (oracle/rt_oracle.c:orc_ray_triangle): whatever
the reference accepts must be a candidate.  Rays: the frame's primary rays (oracle arithmetic), shadow rays from the
oracle's hit points, and shadow rays from random points on the surfaces.  Runs without a GPU.

The second half does the same for the GENERIC filter (every bounce ray, and every ray when the pencil conditions do not
hold): pencil_check.cpp:generic_check restates k_build_records / fast_set / filter_pair with fmaf(); its two approximate
operations (MUFU.RCP, rsqrtf) are perturbed by a few ulp in both directions."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

SO = os.path.join(ROOT, "raytracert_b200", "_build", "pencil_check.so")
SRC = os.path.join(ROOT, "tests", "pencil_check.cpp")
HDR = os.path.join(ROOT, "raytracert_b200", "csrc", "rt_pencil.h")
HDR2 = os.path.join(ROOT, "raytracert_b200", "csrc", "rt_tpencil.h")


class Result(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("pairs", "ref_hits", "candidates", "violations", "grazing_skipped", "unsafe_rays", "always_tris", "never_recs")] + \
               [(n, C.c_int32) for n in ("setup_ok", "first_bad_ray", "first_bad_tri", "pad")] + [("delta", C.c_double), ("M", C.c_double), ("cos_g", C.c_double)]


@pytest.fixture(scope="session")
def checker(port):
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR), os.path.getmtime(HDR2)):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", SO, SRC], check=True)
    L = C.CDLL(SO)
    L.pencil_check.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.POINTER(Result)]
    L.pencil_check_near_planes.restype = C.c_longlong
    L.pencil_check_set_plane.argtypes = [C.c_double] * 4
    pair_fn = C.cast(port.L.orc_ray_triangle, C.c_void_p)

    def run(mode, setup, M, tris, rays, inv_scale=1.0, premise=True):
        """premise=False: no scene-level clause-free proof; triangles whose plane passes within lam_max*cos_g + 2*delta of the
        common point are left to the exact path instead (their number is in run.near_planes afterwards)."""
        setup = np.ascontiguousarray(setup, np.float32)
        tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        r = Result()
        L.pencil_check_set_premise(1 if premise else 0)
        try:
            L.pencil_check(mode, setup.ctypes.data, float(M), len(tris), tris.ctypes.data, len(rays), rays.ctypes.data, inv_scale, pair_fn, C.byref(r))
            run.near_planes = int(L.pencil_check_near_planes())
        finally:
            L.pencil_check_set_premise(1)
        return r
    run.near_planes = 0
    run.set_plane = lambda n, d: L.pencil_check_set_plane(float(n[0]), float(n[1]), float(n[2]), float(d))
    L.tpencil_check.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                C.c_void_p, C.POINTER(Result)]

    def run_tp(eye, delta_cam, M, box_lo, box_hi, tris, rays, ray_tri, R=8):
        eye = np.ascontiguousarray(eye, np.float64); lo = np.ascontiguousarray(box_lo, np.float32); hi = np.ascontiguousarray(box_hi, np.float32)
        tris_ = np.ascontiguousarray(tris, np.float32).reshape(-1, 9); rays_ = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        rt = np.ascontiguousarray(ray_tri, np.int32)
        r = Result()
        L.tpencil_check(eye.ctypes.data, float(delta_cam), float(M), lo.ctypes.data, hi.ctypes.data, len(tris_), tris_.ctypes.data, len(rays_), rays_.ctypes.data,
                        rt.ctypes.data, int(R), pair_fn, C.byref(r))
        return r
    run.thread_pencil = run_tp
    return run


@pytest.fixture(scope="session")
def generic_checker(checker, port):
    """CPU restatement of the GENERIC filter (pencil_check.cpp:generic_check), same calling convention."""
    L = C.CDLL(SO)
    L.generic_check.argtypes = [C.c_int, C.c_double, C.c_float, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.POINTER(Result)]
    pair_fn = C.cast(port.L.orc_ray_triangle, C.c_void_p)

    def run(mode, M, bmin, tris, rays, rc_scale=1.0, inv_scale=1.0):
        tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        r = Result()
        L.generic_check(mode, float(M), bmin, len(tris), tris.ctypes.data, len(rays), rays.ctypes.data, rc_scale, inv_scale, pair_fn, C.byref(r))
        return r
    return run


def pow2_ceil(v):
    m = 1.0 / 1024.0
    while m < v:
        m *= 2.0
    return m


def magnitude_bound(scene, corners):
    """rt_b200.cu:magnitude_bound."""
    ext = float(np.abs(scene.vertices[np.isfinite(scene.vertices)]).max()) if scene.vertices.size else 0.0
    c = np.asarray(corners, np.float32).reshape(4, 6)
    return pow2_ceil(max(ext + 0.5, float(np.abs(c[:, :3]).max())))


def primary_rays(corners, W, H, pf, step):
    """main.cpp:380-386 in float32, every step-th pixel of every step-th row, all sub-samples."""
    c = np.asarray(corners, np.float32).reshape(4, 2, 3)   # c00, c01, c10, c11 x (origin, dest)
    f = np.float32
    divX, divY = f(W * pf - 1), f(H * pf - 1)
    out = []
    for y in range(0, H, step):
        for x in range(0, W, step):
            for sx in range(pf):
                for sy in range(pf):
                    xs = f(1) - (f(x) * f(pf) + f(sx)) / divX
                    ys = f(1) - (f(y) * f(pf) + f(sy)) / divY
                    omx, omy = f(1) - xs, f(1) - ys
                    ray = []
                    for k in range(2):
                        top = (c[0, k] * xs + c[2, k] * omx) * ys
                        bot = (c[1, k] * xs + c[3, k] * omx) * omy
                        ray.append((top + bot).astype(np.float32))
                    out.append(np.concatenate(ray))
    return np.array(out, np.float32)


def tri_array(scene):
    return scene.vertices[scene.indices].reshape(-1, 9)


def scene_box(tris, M):
    """A box at least as large as the union of the tile boxes (k_build_tile_boxes): the triangles + a margin."""
    t = tris.reshape(-1, 3)
    t = t[np.isfinite(t).all(axis=1)]
    m = 0.05 + M * 2.0 ** -14
    return t.min(axis=0) - m, t.max(axis=0) + m


def edge_points(tris, rng, n):
    """Points ON the edges and AT the vertices of random triangles (float32): rays aimed at them make the reference's
    own inside/outside decision a coin flip, which is where a filter tolerance that is too tight would show."""
    t = tris.reshape(-1, 3, 3)
    pick = rng.integers(0, len(t), n)
    e = rng.integers(0, 3, n)
    u = rng.uniform(0, 1, n).astype(np.float32)
    u[: n // 4] = 0.0                                   # vertices
    a, b = t[pick, e], t[pick, (e + 1) % 3]
    return (a + (b - a) * u[:, None]).astype(np.float32)


def grazing_product(tris, rays):
    """max |u||v| over the triangles x max |dest - origin| over the rays (rt_b200.cu:build_records)."""
    t = tris.reshape(-1, 3, 3).astype(np.float64)
    uv = np.linalg.norm(t[:, 1] - t[:, 0], axis=1) * np.linalg.norm(t[:, 2] - t[:, 0], axis=1)
    r = rays.astype(np.float64)
    return float(uv.max() * np.linalg.norm(r[:, 3:] - r[:, :3], axis=1).max())


def check_frame(checker, port, scene, cam, W, H, pf, lights, step, expect_cam=True):
    tris = tri_array(scene)
    M = magnitude_bound(scene, cam.corners)
    rays = primary_rays(cam.corners, W, H, pf, step)
    if expect_cam:
        # adversarial primary rays: from the eye through edge / vertex points, origin one unit from the eye like a
        # near-plane origin (their lines pass within rounding of the eye, far inside the set-up's delta)
        rng = np.random.default_rng(11)
        P = edge_points(tris, rng, 1500)
        eye = np.asarray(cam.eye, np.float32)
        d = P - eye
        d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        fwd = np.asarray(cam.corners, np.float32).reshape(4, 2, 3)
        fwd = (fwd[:, 1] - fwd[:, 0]).mean(axis=0)
        keep = d @ fwd > 0                              # in front of the camera
        adv = np.concatenate([eye + d[keep], eye + np.float32(9.0) * d[keep]], axis=1).astype(np.float32)
    else:
        adv = rays[:0]
    total = dict(pairs=0, ref_hits=0, candidates=0)
    n_lattice = len(rays)
    for scale in (1.0, 1.0 + 3 * 2.0 ** -24, 1.0 - 3 * 2.0 ** -24):     # rsqrtf is accurate to ~2 ulp
        for batch in (rays, adv):
            if not len(batch):
                continue
            r = checker(0, cam.corners, M, tris, batch, scale)
            assert bool(r.setup_ok) == expect_cam, "camera pencil set-up"
            if not r.setup_ok:
                break
            assert r.violations == 0, f"primary: {r.violations} accepted pairs were filtered out (first: ray {r.first_bad_ray}, triangle {r.first_bad_tri})"
            if r.cos_g * grazing_product(tris, batch) <= 0.9e-5:    # the clause-free premise holds for this batch
                assert r.grazing_skipped == 0
            if batch is rays:     # selectivity is judged on the frame's own rays
                for k in total:
                    total[k] += getattr(r, k)
    rays = np.concatenate([rays, adv])
    # shadow rays: from the oracle's hit points of these primary rays, and from random surface points
    port.set_scene(scene)
    port.configure(cam.eye, np.zeros((0, 3), np.float32), 0, 0)
    _, prim, hit = port.trace(rays[:, :3], rays[:, 3:])
    origins = [hit[prim >= 0]]
    rng = np.random.default_rng(5)
    pick = rng.integers(0, len(tris), 400)
    bary = rng.dirichlet((1, 1, 1), 400).astype(np.float32)
    origins.append((tris[pick].reshape(-1, 3, 3) * bary[:, :, None]).sum(axis=1).astype(np.float32))
    origins = (np.concatenate(origins).astype(np.float32) + np.float32(0.1)).astype(np.float32)   # raytracing.cpp:246
    lo, hi = scene_box(tris, M)
    n_light_ok = 0
    for Lp in lights:
        Lp = np.asarray(Lp, np.float32)
        # adversarial shadow rays: origins placed so that the ray to the light passes through an edge / vertex point
        P = edge_points(tris, rng, 1200)
        back = P + (P - Lp) * rng.uniform(0.05, 1.5, (len(P), 1)).astype(np.float32)
        all_o = np.concatenate([origins, back.astype(np.float32)])
        srays = np.concatenate([all_o, np.broadcast_to(Lp, all_o.shape)], axis=1)
        setup = np.concatenate([Lp, lo, hi]).astype(np.float32)
        for scale in (1.0, 1.0 + 3 * 2.0 ** -24):
            r = checker(1, setup, M, tris, srays, scale)
            if not r.setup_ok:
                break
            assert r.violations == 0, f"shadow {Lp}: {r.violations} accepted pairs were filtered out (first: ray {r.first_bad_ray}, triangle {r.first_bad_tri})"
            if r.cos_g * grazing_product(tris, srays) <= 0.9e-5:
                assert r.grazing_skipped == 0
        n_light_ok += int(r.setup_ok)
    return total, n_light_ok


def test_pencil_sound_on_the_balls_standin(checker, port):
    """Reduced Balls stand-in (same generator as the headline scene), default and low grazing camera, near lights."""
    from raytracert_b200 import host, scenes
    s = scenes.balls_standin(grid=40, slices=24, stacks=12)
    cam = host.Camera(64, 48)
    total, nl = check_frame(checker, port, s, cam, 64, 48, 2, [tuple(cam.eye), (2.5, 4.0, 3.0)], step=3)
    assert total["ref_hits"] > 500 and nl == 2
    # selectivity: the filter must stay a filter (a few candidates per accepted pair, not "everything")
    assert total["candidates"] < 20 * total["ref_hits"]
    cam2 = host.Camera(64, 48, (0.2, 0.75, 4.6), (0.0, 0.62, 0.0))     # near-tangent rays over the terrain
    total2, _ = check_frame(checker, port, s, cam2, 64, 48, 2, [(2.5, 4.0, 3.0), (-3.0, 2.0, 0.5)], step=3)
    assert total2["ref_hits"] > 500


def test_pencil_sound_on_a_tessellated_sphere(checker, port):
    from raytracert_b200 import host, scenes
    s = scenes.tessellated_sphere(slices=96, stacks=49)
    cam = host.Camera(48, 48, (1.2, 0.9, 2.6), (0.0, 0.0, 0.0))
    total, nl = check_frame(checker, port, s, cam, 48, 48, 2, [tuple(cam.eye), (0.0, 3.0, 0.0)], step=2)
    assert total["ref_hits"] > 1000 and nl == 2


@pytest.mark.parametrize("seed", range(6))
def test_pencil_sound_on_random_fine_soups(checker, port, seed):
    """The fuzz test's 'fine mesh' kind: random small triangles, random cameras (sometimes inside the soup) and lights."""
    from raytracert_b200 import host
    rng = np.random.default_rng(7000 + seed)
    n = int(rng.integers(50, 400))
    ctr = rng.uniform(-1.5, 1.5, (n, 1, 3))
    tri = ctr + 0.08 * rng.normal(size=(n, 3, 3))
    if seed == 3:
        tri += np.array([40.0, -25.0, 10.0])     # far from the world origin: larger magnitude bound
    v = tri.reshape(-1, 3).astype(np.float32)
    idx = np.arange(len(v), dtype=np.uint32).reshape(-1, 3)
    mats = np.zeros((1, 16), np.float32)
    s = host.Scene(v, idx, np.zeros(n, np.uint32), host.face_normals(v, idx), mats)
    centre = v.mean(axis=0)
    eye = centre + (rng.uniform(-1, 1, 3) + np.array([0, 0, 5.0]) if seed % 3 else rng.uniform(-0.5, 0.5, 3))
    cam = host.Camera(40, 32, tuple(eye), tuple(centre + rng.uniform(-0.5, 0.5, 3)))
    lights = centre + rng.uniform(-6, 6, (3, 3))
    total, _ = check_frame(checker, port, s, cam, 40, 32, 2, [tuple(l) for l in lights], step=1)
    assert total["pairs"] > 0


def _grid_patch(n, size, centre, tilt=0.3):
    """A tilted, slightly bumpy n x n grid patch (2 n^2 triangles) of edge size/n around `centre`."""
    from raytracert_b200 import host
    u = np.linspace(-0.5, 0.5, n + 1)
    X, Z = np.meshgrid(u * size, u * size)
    Y = tilt * X + 0.05 * size * np.sin(9 * X / size) * np.cos(7 * Z / size)
    v = (np.stack([X, Y, Z], axis=-1).reshape(-1, 3) + np.asarray(centre)).astype(np.float32)
    idx = []
    for i in range(n):
        for j in range(n):
            a = i * (n + 1) + j
            idx += [(a, a + 1, a + n + 1), (a + 1, a + n + 2, a + n + 1)]
    idx = np.array(idx, np.uint32)
    return host.Scene(v, idx, np.zeros(len(idx), np.uint32), host.face_normals(v, idx), np.zeros((1, 16), np.float32))


@pytest.mark.parametrize("case", ["wide_frame", "light_close_to_the_box", "tiny_triangles", "far_from_the_origin", "camera_far_away"])
def test_pencil_sound_in_extreme_setups(checker, port, case):
    """Corners of the launch conditions: a very wide frame (large chart extent), a light just far enough from the scene box
    (w_max close to its cap), millimetre triangles (large barycentric gradients), a scene at |coordinate| ~ 60 (large
    magnitude bound), a camera 9 units away (rays near the far plane)."""
    from raytracert_b200 import host
    if case == "wide_frame":
        s = _grid_patch(24, 3.0, (0, 0, 0))
        cam = host.Camera(160, 36, (0.0, 1.2, 3.0), (0.0, 0.0, 0.0))
        lights, W, H = [(0.5, 3.0, 1.0)], 160, 36
    elif case == "light_close_to_the_box":
        s = _grid_patch(24, 3.0, (0, 0, 0))
        cam = host.Camera(48, 36, (0.0, 2.2, 3.4), (0.0, 0.0, 0.0))
        lights, W, H = [(0.3, 1.25, 0.2), (1.9, 1.3, -1.7)], 48, 36      # low over the patch: wide cones
    elif case == "tiny_triangles":
        s = _grid_patch(40, 0.08, (0.2, 0.1, -0.3))
        cam = host.Camera(48, 36, (0.2, 0.6, 0.9), (0.2, 0.1, -0.3))       # the patch covers a few pixels; the edge-aimed rays do the work
        lights, W, H = [(0.4, 1.5, 0.3)], 48, 36
    elif case == "far_from_the_origin":
        s = _grid_patch(24, 3.0, (55.0, -40.0, 20.0))
        cam = host.Camera(48, 36, (55.0, -38.5, 23.5), (55.0, -40.0, 20.0))
        lights, W, H = [(56.0, -36.5, 21.0)], 48, 36
    else:
        s = _grid_patch(24, 3.0, (0, 0, 0))
        cam = host.Camera(48, 36, (0.5, 4.0, 8.0), (0.0, 0.0, 0.0))
        lights, W, H = [(0.0, 6.0, 0.0)], 48, 36
    total, nl = check_frame(checker, port, s, cam, W, H, 2, lights, step=1 if case == "tiny_triangles" else 2)
    assert total["ref_hits"] > (10 if case == "tiny_triangles" else 100), case
    assert nl >= 1, case


@pytest.mark.parametrize("name", ["cube", "cube_default_camera", "shadow_test", "dodge", "mirror_room"])
def test_pencil_sound_without_the_clause_free_premise(checker, port, name):
    """Groundwork for scenes with large triangles (experimental in the library): without the scene-level
    proof that the reference rejects grazing pairs, the pencil filter is still sound for every triangle whose plane stays
    lam_max*cos_g + 2*delta away from the common point -- a pair that can hit inside the scene then has |cos| >= cos_g by
    geometry alone.  The few triangles nearer than that get "always candidate" records (cube under the default camera: the four
    triangles of the two faces whose planes contain the eye).  RT_OPT_PENCIL_ANY wires this into the library (default off)."""
    from conftest import load_scene
    from raytracert_b200 import host, scenes
    s = scenes.mirror_room(n=12) if name == "mirror_room" else load_scene("cube" if name.startswith("cube") else name)
    cam = {"cube": host.Camera(40, 40, (2.6, 2.4, 3.0), (.5, .5, .5)), "cube_default_camera": host.Camera(40, 40),
           "shadow_test": host.Camera(40, 30, (1, 5, 7), (1, 1.2, 0.7)), "dodge": host.Camera(48, 27, (.75, .55, 1.1), (.07, 0, .23)),
           "mirror_room": host.Camera(40, 30, (0.3, 1.6, 4.2), (0, 0.8, 0))}[name]
    tris = tri_array(s)
    M = magnitude_bound(s, cam.corners)
    rng = np.random.default_rng(11)
    rays = primary_rays(cam.corners, cam.W, cam.H, 2, 1)
    P = edge_points(tris, rng, 1500)
    eye = np.asarray(cam.eye, np.float32)
    d = P - eye
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    fwd = np.asarray(cam.corners, np.float32).reshape(4, 2, 3)
    keep = d @ (fwd[:, 1] - fwd[:, 0]).mean(axis=0) > 0
    adv = np.concatenate([eye + d[keep], eye + np.float32(9.0) * d[keep]], axis=1).astype(np.float32)
    accepted = 0
    for batch in (rays, adv):
        r = checker(0, cam.corners, M, tris, batch, premise=False)
        assert r.setup_ok and r.violations == 0 and r.grazing_skipped == 0, f"{name}: {r.violations} accepted pairs were filtered out"
        lo_n, hi_n = {"cube_default_camera": (4, 4), "dodge": (1, 4)}.get(name, (0, 0))
        assert lo_n <= checker.near_planes <= hi_n
        accepted += r.ref_hits
    assert accepted > 1000
    port.set_scene(s)
    port.configure(cam.eye, np.zeros((0, 3), np.float32), 0, 0)
    _, prim, hit = port.trace(rays[:, :3], rays[:, 3:])
    lo, hi = scene_box(tris, M)
    for Lp in (np.asarray(cam.eye, np.float32), np.array([3.0, 6.0, 2.0], np.float32)):
        Pe = edge_points(tris, rng, 800)
        back = (Pe + (Pe - Lp) * rng.uniform(0.05, 1.5, (len(Pe), 1)).astype(np.float32)).astype(np.float32)
        o = np.concatenate([(hit[prim >= 0] + np.float32(0.1)).astype(np.float32), back])
        r = checker(1, np.concatenate([Lp, lo, hi]).astype(np.float32), M, tris, np.concatenate([o, np.broadcast_to(Lp, o.shape)], axis=1), premise=False)
        if r.setup_ok:
            assert r.violations == 0 and r.grazing_skipped == 0, f"{name}, light {Lp}: {r.violations} accepted pairs were filtered out"


def test_reflection_pencil_around_the_mirrored_eye(checker, port):
    """Groundwork: the reflections of primary rays off one plane (here the stand-in's water, y = 0) form a
    pencil through the mirror image of the eye.  Reflected rays computed like reflection() / addOffset()
    (raytracing.cpp:266-285) in float32; the same records and the same filter, around E*, must keep every accepted pair."""
    from raytracert_b200 import host, scenes
    f = np.float32
    s = scenes.balls_standin(grid=48, slices=24, stacks=12)
    cam = host.Camera(120, 120, (0.0, 2.6, 5.2), (0.0, 0.55, 0.0))
    tris = tri_array(s)
    M = magnitude_bound(s, cam.corners)
    rays = primary_rays(cam.corners, 120, 120, 2, 1)
    port.set_scene(s)
    port.configure(cam.eye, np.zeros((0, 3), np.float32), 0, 0)
    _, prim, hit = port.trace(rays[:, :3], rays[:, 3:])
    nrm = s.normals[np.maximum(prim, 0)]
    flat = np.all(tris.reshape(-1, 3, 3)[:, :, 1] == 0.0, axis=1)          # triangles lying exactly in the plane y = 0
    sel = np.where((prim >= 0) & flat[np.maximum(prim, 0)])[0]
    assert len(sel) > 2000
    P = hit[sel].astype(f)
    ray = (rays[sel, 3:] - rays[sel, :3]).astype(f)
    r = (ray * (f(1) / np.sqrt((ray * ray).sum(1, dtype=f)).astype(f))[:, None]).astype(f)
    nf = nrm[sel].astype(f)
    R = (r - (f(2) * (nf * r).sum(1, dtype=f))[:, None] * nf).astype(f)
    dest = (P + R).astype(f)
    off = (dest - P).astype(f)
    off = (off * (f(1) / np.sqrt((off * off).sum(1, dtype=f)).astype(f))[:, None]).astype(f)
    point = (P + off * f(0.01)).astype(f)
    brays = np.concatenate([point, dest], axis=1).astype(f)
    estar = np.asarray(cam.eye, np.float64) * np.array([1.0, -1.0, 1.0])
    dd = (dest - point).astype(np.float64)
    dd /= np.linalg.norm(dd, axis=1, keepdims=True)
    w = estar - point.astype(np.float64)
    delta = np.linalg.norm(w - (w * dd).sum(1)[:, None] * dd, axis=1).max()
    assert delta < 1e-5                                  # the camera pencil's own delta is 5e-6
    fax = dd.mean(axis=0)
    fax /= np.linalg.norm(fax)
    setup = np.zeros(24, np.float32)
    setup[:3], setup[3:6] = estar, fax
    setup[6], setup[7], setup[8] = 4 * delta, (1.0 / (dd @ fax)).max() * 1.05, np.linalg.norm(point - estar, axis=1).min() * 0.9
    for scale in (1.0, 1.0 + 3 * 2.0 ** -24):
        res = checker(2, setup, M, tris, brays, scale)
        assert res.setup_ok and res.violations == 0 and res.grazing_skipped == 0 and res.unsafe_rays == 0
    assert res.candidates < 3 * len(brays)              # about one per ray: the triangle it starts on


def reflected_rays(rays, hit, normals):
    """Continuation rays exactly as reflection() / addOffset() build them (raytracing.cpp:266-285), in float32."""
    f = np.float32
    P = hit.astype(f)
    ray = (rays[:, 3:] - rays[:, :3]).astype(f)
    r = (ray * (f(1) / np.sqrt((ray * ray).sum(1, dtype=f)).astype(f))[:, None]).astype(f)
    nf = normals.astype(f)
    R = (r - (f(2) * (nf * r).sum(1, dtype=f))[:, None] * nf).astype(f)
    dest = (P + R).astype(f)
    off = (dest - P).astype(f)
    off = (off * (f(1) / np.sqrt((off * off).sum(1, dtype=f)).astype(f))[:, None]).astype(f)
    return np.concatenate([(P + off * f(0.01)).astype(f), dest], axis=1).astype(f)


@pytest.mark.parametrize("case", ["water_default", "water_low_camera", "room_floor", "room_wall", "cube_face"])
def test_shipped_mirror_pencil_is_sound(checker, port, case):
    """RT_OPT_PENCIL_REFLECT as shipped: pencil_mirror_setup() (camera pencil mirrored about a plane group) + the runtime check
    pencil_mirror_accepts() k_shade applies to every continuation ray.  Level-1 continuation rays of the primary hits -- on the
    plane AND elsewhere (those must be refused by the check or be harmless) -- go through the records around the mirrored eye;
    no pair the reference accepts may be filtered out; rays reflected off the plane itself must (almost all) be accepted."""
    from raytracert_b200 import host, scenes
    if case.startswith("water"):
        s = scenes.balls_standin(grid=48, slices=24, stacks=12)
        cam = host.Camera(96, 96, (0.0, 2.6, 5.2), (0.0, 0.55, 0.0)) if case == "water_default" else host.Camera(96, 64, (0.2, 0.75, 4.6), (0.0, 0.62, 0.0))
        n, d, premise = (0.0, 1.0, 0.0), 0.0, True
    elif case.startswith("room"):
        s = scenes.mirror_room(n=12)
        cam = host.Camera(64, 48, (0.3, 1.6, 4.2), (0, 0.8, 0))
        n, d, premise = ((0.0, 1.0, 0.0), 0.0, False) if case == "room_floor" else ((0.0, 0.0, 1.0), float(s.vertices[:, 2].min()), False)
    else:
        from conftest import load_scene
        s = load_scene("cube")
        cam = host.Camera(64, 64, (2.6, 2.4, 3.0), (.5, .5, .5))
        n, d, premise = (0.0, 1.0, 0.0), 1.0, False
    tris = tri_array(s)
    M = magnitude_bound(s, cam.corners)
    rays = primary_rays(cam.corners, cam.W, cam.H, 2, 1)
    port.set_scene(s)
    port.configure(cam.eye, np.zeros((0, 3), np.float32), 0, 0)
    _, prim, hit = port.trace(rays[:, :3], rays[:, 3:])
    ok = prim >= 0
    brays = reflected_rays(rays[ok], hit[ok], s.normals[prim[ok]])
    t = tris.reshape(-1, 3, 3)[prim[ok]]
    on_plane = np.all(np.abs(t @ np.asarray(n, np.float64) - d) < 1e-6, axis=1)
    checker.set_plane(n, d)
    try:
        for scale in (1.0, 1.0 + 3 * 2.0 ** -24):
            res = checker(3, cam.corners, M, tris, brays, scale, premise=premise)
            assert res.setup_ok, case
            assert res.violations == 0, f"{case}: {res.violations} accepted pairs filtered out (ray {res.first_bad_ray}, triangle {res.first_bad_tri})"
        if on_plane.sum() > 50:
            res_on = checker(3, cam.corners, M, tris, brays[on_plane], 1.0, premise=premise)
            assert res_on.unsafe_rays <= 0.02 * on_plane.sum(), f"{case}: the check refuses {res_on.unsafe_rays} of {on_plane.sum()} rays reflected off the plane"
            assert res_on.violations == 0
        if (~on_plane).sum() > 50:      # rays off other surfaces: nearly all must be refused (they do not pass through E*)
            res_off = checker(3, cam.corners, M, tris, brays[~on_plane], 1.0, premise=premise)
            assert res_off.violations == 0
    finally:
        checker.set_plane((0, 1, 0), 0)


@pytest.mark.parametrize("case", ["balls", "balls_low_camera", "sphere", "room", "cube", "dodge"])
def test_thread_pencils_are_sound(checker, port, case):
    """Thread pencils (rt_tpencil.h, RT_OPT_PENCIL_THREAD): the level-1 continuation rays of the primary hits on ONE triangle share
    the mirror image of the eye about that triangle's plane.  Replay of what the kernels do -- mirror point, acceptance check,
    groups of 8 accepted rays per reflector, record + on-the-fly orientation, hot test, full test with the tightest admissible
    distance bound -- for EVERY (continuation ray, triangle) pair against the oracle: no accepted pair may be filtered out, most
    rays must be accepted, and the hot test must stay selective."""
    from conftest import load_scene
    from raytracert_b200 import host, scenes
    if case.startswith("balls"):
        s = scenes.balls_standin(grid=48, slices=24, stacks=12)
        cam = host.Camera(72, 72, (0.0, 2.6, 5.2), (0.0, 0.55, 0.0)) if case == "balls" else host.Camera(72, 48, (0.2, 0.75, 4.6), (0.0, 0.62, 0.0))
    elif case == "sphere":
        s = scenes.tessellated_sphere(slices=96, stacks=49, ground=True)
        cam = host.Camera(64, 64, (1.2, 0.9, 2.6), (0.0, 0.0, 0.0))
    elif case == "room":
        s = scenes.mirror_room(n=12)
        cam = host.Camera(56, 42, (0.3, 1.6, 4.2), (0, 0.8, 0))
    elif case == "cube":
        s = load_scene("cube")
        cam = host.Camera(48, 48, (2.6, 2.4, 3.0), (.5, .5, .5))
    else:
        s = load_scene("dodge")
        cam = host.Camera(48, 27, (.75, .55, 1.1), (.07, 0, .23))
    tris = tri_array(s)
    M = magnitude_bound(s, cam.corners)
    rays = primary_rays(cam.corners, cam.W, cam.H, 3, 1)        # 9 samples per pixel: several rays per facet
    port.set_scene(s)
    port.configure(cam.eye, np.zeros((0, 3), np.float32), 0, 0)
    _, prim, hit = port.trace(rays[:, :3], rays[:, 3:])
    ok = prim >= 0
    brays = reflected_rays(rays[ok], hit[ok], s.normals[prim[ok]])
    t = tris.reshape(-1, 9)
    lo = t.reshape(-1, 3).min(axis=0) - 0.01; hi = t.reshape(-1, 3).max(axis=0) + 0.01
    # the camera pencil's centre and delta, as plan_pencil has them (the eye itself is within ~1e-6 of it)
    res = checker.thread_pencil(np.asarray(cam.eye, np.float64), 6e-6, M, lo, hi, tris, brays, prim[ok])
    assert res.setup_ok, case
    assert res.violations == 0, f"{case}: {res.violations} accepted pairs filtered out (ray {res.first_bad_ray}, triangle {res.first_bad_tri})"
    n = len(brays)
    assert res.unsafe_rays <= 0.05 * n, f"{case}: the acceptance check refuses {res.unsafe_rays} of {n} continuation rays"
    assert res.ref_hits > 0 or case == "cube"        # (nothing to hit around a single convex cube)
    # selectivity: hot candidates (cold-path entries) and what survives the full test (exact evaluations at most), per accepted ray
    acc = n - res.unsafe_rays
    print(f"{case}: {n} rays, {res.unsafe_rays} refused, hot candidates/ray {res.candidates / acc:.2f}, full-test survivors/ray {res.grazing_skipped / acc:.2f}, "
          f"reference hits/ray {res.ref_hits / acc:.2f}, delta {res.delta:.2e}")
    assert res.candidates <= 40 * acc and res.grazing_skipped <= 6 * acc


@pytest.mark.parametrize("seed", range(int(os.environ.get("RT_TP_SEEDS", "8"))))
def test_thread_pencils_sound_on_random_soups(checker, port, seed):
    """Random triangle soups -- small and large triangles, slivers (seed % 4 == 1), a soup far from the world origin (seed 3), a
    camera inside the soup (seed % 3 == 0) -- through the thread-pencil replay: primary rays hit whatever they hit, the
    continuation rays are grouped by reflector.  Every pair the reference accepts must survive; rays the check refuses only cost
    speed."""
    from raytracert_b200 import host
    rng = np.random.default_rng(9100 + seed)
    n = int(rng.integers(60, 350))
    ctr = rng.uniform(-1.5, 1.5, (n, 1, 3))
    tri = ctr + rng.choice([0.05, 0.3, 0.9]) * rng.normal(size=(n, 3, 3))
    if seed % 4 == 1:      # slivers: the third vertex almost on the line of the first two
        t = rng.uniform(-0.2, 1.2, (n, 1))
        tri[:, 2] = tri[:, 0] + t * (tri[:, 1] - tri[:, 0]) + 1e-4 * rng.normal(size=(n, 3))
    if seed == 3:
        tri += np.array([40.0, -25.0, 10.0])
    v = tri.reshape(-1, 3).astype(np.float32)
    idx = np.arange(len(v), dtype=np.uint32).reshape(-1, 3)
    s = host.Scene(v, idx, np.zeros(n, np.uint32), host.face_normals(v, idx), np.zeros((1, 16), np.float32))
    centre = v.mean(axis=0)
    eye = centre + (rng.uniform(-1, 1, 3) + np.array([0, 0, 5.0]) if seed % 3 else rng.uniform(-0.5, 0.5, 3))
    cam = host.Camera(40, 32, tuple(eye), tuple(centre + rng.uniform(-0.5, 0.5, 3)))
    tris = tri_array(s)
    M = magnitude_bound(s, cam.corners)
    rays = primary_rays(cam.corners, 40, 32, 3, 1)
    port.set_scene(s)
    port.configure(cam.eye, np.zeros((0, 3), np.float32), 0, 0)
    _, prim, hit = port.trace(rays[:, :3], rays[:, 3:])
    ok = prim >= 0
    if ok.sum() == 0:      # (slivers are hard to see; use a denser frame for them below)
        rays = primary_rays(cam.corners, 40, 32, 9, 1)
        _, prim, hit = port.trace(rays[:, :3], rays[:, 3:])
        ok = prim >= 0
    if ok.sum() == 0:
        return             # nothing visible even then: nothing to check
    brays = reflected_rays(rays[ok], hit[ok], s.normals[prim[ok]])
    t3 = tris.reshape(-1, 3)
    res = checker.thread_pencil(np.asarray(cam.eye, np.float64), 1e-5, M, t3.min(axis=0) - 0.01, t3.max(axis=0) + 0.01, tris, brays, prim[ok])
    assert res.setup_ok
    assert res.violations == 0, f"seed {seed}: {res.violations} accepted pairs filtered out (ray {res.first_bad_ray}, triangle {res.first_bad_tri})"
    assert res.pairs > 0 or res.unsafe_rays == len(brays)


def test_thread_pencil_replay_detects_a_broken_filter(checker, port, tmp_path):
    """The replay must be able to fail: the same harness built from a header in which tp_orient() no longer orients the weight
    vectors by the side of the plane the common point lies on reports thousands of filtered-out accepted pairs."""
    from raytracert_b200 import host, scenes
    src = open(HDR2).read()
    assert "    m &= 0x80000000u;" in src
    mut = tmp_path / "raytracert_b200" / "csrc"
    mut.mkdir(parents=True)
    (mut / "rt_tpencil.h").write_text(src.replace("    m &= 0x80000000u;", "    m = 0u;"))
    (mut / "rt_pencil.h").write_text(open(HDR).read())
    (tmp_path / "tests").mkdir()
    (tmp_path / "tests" / "pencil_check.cpp").write_text(open(SRC).read())
    so = str(tmp_path / "pc_mut.so")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", so, str(tmp_path / "tests" / "pencil_check.cpp")], check=True)
    L = C.CDLL(so)
    L.tpencil_check.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                C.c_void_p, C.POINTER(Result)]
    s = scenes.mirror_room(n=12)
    cam = host.Camera(40, 30, (0.3, 1.6, 4.2), (0, 0.8, 0))
    tris = np.ascontiguousarray(tri_array(s), np.float32).reshape(-1, 9)
    M = magnitude_bound(s, cam.corners)
    rays = primary_rays(cam.corners, 40, 30, 3, 1)
    port.set_scene(s)
    port.configure(cam.eye, np.zeros((0, 3), np.float32), 0, 0)
    _, prim, hit = port.trace(rays[:, :3], rays[:, 3:])
    ok = prim >= 0
    brays = np.ascontiguousarray(reflected_rays(rays[ok], hit[ok], s.normals[prim[ok]]), np.float32)
    refl = np.ascontiguousarray(prim[ok], np.int32)
    t3 = tris.reshape(-1, 3)
    lo = (t3.min(axis=0) - 0.01).astype(np.float32); hi = (t3.max(axis=0) + 0.01).astype(np.float32)
    eye = np.ascontiguousarray(cam.eye, np.float64)
    r = Result()
    L.tpencil_check(eye.ctypes.data, 6e-6, float(M), lo.ctypes.data, hi.ctypes.data, len(tris), tris.ctypes.data, len(brays), brays.ctypes.data, refl.ctypes.data, 8,
                    C.cast(port.L.orc_ray_triangle, C.c_void_p), C.byref(r))
    good = checker.thread_pencil(eye, 6e-6, M, lo, hi, tris, brays, refl)
    assert good.violations == 0 and good.ref_hits > 1000
    assert r.violations > 0.5 * good.ref_hits, (r.violations, good.ref_hits)


@pytest.mark.parametrize("seed", range(int(os.environ.get("RT_MIRROR_SEEDS", "6"))))
def test_mirror_pencil_on_random_tilted_planes(checker, port, seed):
    """A tilted planar patch (float32 vertices: coplanar only up to rounding -- the group's plane is its first triangle's, as in
    rt_upload_scene) under random blobs, random cameras.  Continuation rays of the primary hits anywhere go through the mirror
    pencil of that plane: whatever the runtime check accepts must keep every pair the reference accepts; most rays off the
    patch must be accepted."""
    from raytracert_b200 import host
    rng = np.random.default_rng(4200 + seed)
    nq = int(rng.integers(6, 20))
    a, b, c = rng.uniform(-0.6, 0.6), rng.uniform(-0.6, 0.6), rng.uniform(-0.5, 0.5)
    u = np.linspace(-2.0, 2.0, nq + 1)
    X, Z = np.meshgrid(u, u)
    P = np.stack([X, a * X + b * Z + c, Z], axis=-1).reshape(-1, 3).astype(np.float32)
    ii, jj = np.meshgrid(np.arange(nq), np.arange(nq))
    q = (jj * (nq + 1) + ii).ravel()
    patch = np.concatenate([np.stack([q, q + nq + 1, q + 1], 1), np.stack([q + 1, q + nq + 1, q + nq + 2], 1)])
    nb = int(rng.integers(30, 150))
    ctr = rng.uniform(-1.5, 1.5, (nb, 1, 3)) + np.array([0, 1.2, 0])
    blobs = (ctr + 0.15 * rng.normal(size=(nb, 3, 3))).reshape(-1, 3).astype(np.float32)
    v = np.concatenate([P, blobs])
    idx = np.concatenate([patch, len(P) + np.arange(3 * nb).reshape(-1, 3)]).astype(np.uint32)
    s = host.Scene(v, idx, np.zeros(len(idx), np.uint32), host.face_normals(v, idx), np.zeros((1, 16), np.float32))
    eye = np.array([rng.uniform(-1, 1), rng.uniform(1.5, 3.5), rng.uniform(3.0, 5.0)])
    cam = host.Camera(48, 36, tuple(eye), (rng.uniform(-0.5, 0.5), rng.uniform(0.0, 0.6), rng.uniform(-0.5, 0.5)))
    tris = tri_array(s)
    M = magnitude_bound(s, cam.corners)
    rays = primary_rays(cam.corners, 48, 36, 2, 1)
    port.set_scene(s)
    port.configure(cam.eye, np.zeros((0, 3), np.float32), 0, 0)
    _, prim, hit = port.trace(rays[:, :3], rays[:, 3:])
    ok = prim >= 0
    brays = reflected_rays(rays[ok], hit[ok], s.normals[prim[ok]])
    on_patch = prim[ok] < len(patch)
    # the group's plane: its first triangle's, in double (rt_upload_scene)
    t0 = tris.reshape(-1, 3, 3)[0].astype(np.float64)
    n = np.cross(t0[1] - t0[0], t0[2] - t0[0]); n /= np.linalg.norm(n)
    checker.set_plane(n, float(n @ t0[0]))
    try:
        res = checker(3, cam.corners, M, tris, brays, 1.0, premise=False)
        if not res.setup_ok:
            return         # eye too close to the plane / frame too wide: the library serves no mirror pencil either
        assert res.violations == 0, f"seed {seed}: {res.violations} accepted pairs filtered out (ray {res.first_bad_ray}, triangle {res.first_bad_tri})"
        if on_patch.sum() > 100:
            r_on = checker(3, cam.corners, M, tris, brays[on_patch], 1.0, premise=False)
            assert r_on.unsafe_rays <= 0.1 * on_patch.sum(), f"seed {seed}: {r_on.unsafe_rays} of {on_patch.sum()} rays off the patch refused"
    finally:
        checker.set_plane((0, 1, 0), 0)


def bounce_like_rays(tris, rng, n):
    """Continuation-ray shaped rays (raytracing.cpp:266-285): origin = P + 0.01 * dir, dest = P + dir, P on a surface; half of
    them aimed at an edge / vertex point of another triangle."""
    t = tris.reshape(-1, 3, 3)
    pick = rng.integers(0, len(t), n)
    bary = rng.dirichlet((1, 1, 1), n).astype(np.float32)
    P = (t[pick] * bary[:, :, None]).sum(axis=1).astype(np.float32)
    d = rng.normal(size=(n, 3))
    target = edge_points(tris, rng, n)
    d[: n // 2] = (target - P)[: n // 2]
    d = (d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-12)).astype(np.float32)
    return np.concatenate([P + np.float32(0.01) * d, P + d], axis=1).astype(np.float32)


def check_generic(generic_checker, port, scene, cam, W, H, pf, lights, step, clause_free):
    tris = tri_array(scene)
    M = magnitude_bound(scene, cam.corners)
    rng = np.random.default_rng(17)
    rays = primary_rays(cam.corners, W, H, pf, step)
    port.set_scene(scene)
    port.configure(cam.eye, np.zeros((0, 3), np.float32), 0, 0)
    _, prim, hit = port.trace(rays[:, :3], rays[:, 3:])
    origins = (hit[prim >= 0] + np.float32(0.1)).astype(np.float32)
    batches = [("primary", rays), ("bounce", bounce_like_rays(tris, rng, 1500))]
    for Lp in lights:
        Lp = np.asarray(Lp, np.float32)
        P = edge_points(tris, rng, 600)
        back = (P + (P - Lp) * rng.uniform(0.05, 1.5, (len(P), 1)).astype(np.float32)).astype(np.float32)
        o = np.concatenate([origins, back])
        batches.append(("shadow", np.concatenate([o, np.broadcast_to(Lp, o.shape)], axis=1)))
    hits = 0
    for name, batch in batches:
        bmin = 1.0e-5
        if clause_free and grazing_product(tris, batch) * 1.11e-5 < 1e-5:     # rt_b200.cu:build_records
            bmin = -0.5
        for mode in (0, 1):
            for rc_scale, inv_scale in ((1.0, 1.0), (1.0 + 2.0 ** -22, 1.0 - 2.0 ** -22), (1.0 - 2.0 ** -22, 1.0 + 2.0 ** -22)):
                r = generic_checker(mode, M, bmin, tris, batch, rc_scale, inv_scale)
                assert r.violations == 0, f"{name} rays, mode {mode}: {r.violations} accepted pairs were filtered out (first: ray {r.first_bad_ray}, triangle {r.first_bad_tri})"
        hits += r.ref_hits
    return hits


def test_generic_filter_sound_on_the_balls_standin(generic_checker, port):
    """The same all-pairs check for the generic filter (k_trace / k_shadow without the pencil option; every bounce ray)."""
    from raytracert_b200 import host, scenes
    s = scenes.balls_standin(grid=40, slices=24, stacks=12)
    cam = host.Camera(64, 48, (0.2, 0.75, 4.6), (0.0, 0.62, 0.0))
    assert check_generic(generic_checker, port, s, cam, 64, 48, 2, [(2.5, 4.0, 3.0)], 3, clause_free=True) > 500


@pytest.mark.parametrize("name", ["cube", "shadow_test", "dodge"])
def test_generic_filter_sound_on_reference_scenes(generic_checker, port, name):
    """Scenes shipped with the reference (large triangles, slivers: the grazing clause stays)."""
    from conftest import load_scene
    from raytracert_b200 import host
    s = load_scene(name)
    cam = {"cube": host.Camera(40, 40, (2.6, 2.4, 3.0), (.5, .5, .5)), "shadow_test": host.Camera(40, 30, (1, 5, 7), (1, 1.2, 0.7)),
           "dodge": host.Camera(48, 27, (.75, .55, 1.1), (.07, 0, .23))}[name]
    lights = [tuple(cam.eye)]
    assert check_generic(generic_checker, port, s, cam, cam.W, cam.H, 1, lights, 2 if name == "dodge" else 1, clause_free=False) > 50


@pytest.mark.parametrize("seed", range(4))
def test_generic_filter_sound_on_random_soups(generic_checker, port, seed):
    """Random soups with triangle sizes over two decades, slivers included (the fuzz test's kinds 0 and 3)."""
    from raytracert_b200 import host
    rng = np.random.default_rng(9000 + seed)
    n = int(rng.integers(50, 300))
    ctr = rng.uniform(-1.5, 1.5, (n, 1, 3))
    size = 10 ** rng.uniform(-2.0, 0.2, (n, 1, 1))
    tri = ctr + size * rng.normal(size=(n, 3, 3))
    if seed % 2:
        tri[1, 2] = tri[1, 0] + (tri[1, 1] - tri[1, 0]) * 0.5          # exactly collinear in double, not in float
        tri[2, 2] = tri[2, 1] + 1e-6 * (tri[2, 0] - tri[2, 1])         # sliver
    v = tri.reshape(-1, 3).astype(np.float32)
    idx = np.arange(len(v), dtype=np.uint32).reshape(-1, 3)
    s = host.Scene(v, idx, np.zeros(n, np.uint32), host.face_normals(v, idx), np.zeros((1, 16), np.float32))
    cam = host.Camera(40, 32, tuple(rng.uniform(-1, 1, 3) + np.array([0, 0, 5.0])), tuple(rng.uniform(-0.5, 0.5, 3)))
    lights = rng.uniform(-6, 6, (2, 3))
    check_generic(generic_checker, port, s, cam, 40, 32, 1, [tuple(l) for l in lights], 1, clause_free=False)


def test_camera_setup_rejects_what_is_not_a_pencil(checker):
    """Parallel (orthographic) corner rays, or an eye between the ray origins and the scene, must not use the pencil."""
    from raytracert_b200 import host
    tris = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32)
    rays = np.array([[0, 0, 4, 0, 0, -6]], np.float32)
    ortho = np.array([[-1, 1, 4, -1, 1, -6], [-1, -1, 4, -1, -1, -6], [1, 1, 4, 1, 1, -6], [1, -1, 4, 1, -1, -6]], np.float32)
    assert not checker(0, ortho.reshape(-1), 8.0, tris, rays).setup_ok
    cam = host.Camera(32, 32)
    c = cam.corners.reshape(4, 2, 3).copy()
    c = c[:, ::-1, :]                                   # origin and dest swapped: rays run TOWARDS the common point
    assert not checker(0, np.ascontiguousarray(c).reshape(-1), 8.0, tris, rays).setup_ok
    assert checker(0, cam.corners, 8.0, tris, rays).setup_ok
    # light inside the scene box: no pencil for its shadow rays
    setup = np.array([0.2, 0.2, 0.0, -1, -1, -1, 1, 1, 1], np.float32)
    assert not checker(1, setup, 8.0, tris, rays).setup_ok
    setup[:3] = (0.2, 0.2, 3.0)
    assert checker(1, setup, 8.0, tris, rays).setup_ok
