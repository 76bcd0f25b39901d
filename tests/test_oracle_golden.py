"""The oracle is pinned here: the plain-C restatement (oracle/rt_oracle.c) must reproduce, bit for bit, the
fixtures that the UNMODIFIED reference (oracle/_ref) generated (tools/make_golden.py); where oracle/_ref is
present it is re-run too.  The reference ships no golden vectors of its own."""
import numpy as np
import pytest

from conftest import bits, load_case, load_scene, render_cases, GOLDEN


def _render(O, c, want_samples=True, **kw):
    O.set_scene(load_scene(c["scene"]))
    O.configure(c["eye"], c["lights"], c["features"], c["max_lvl"])
    return O.render(c["corners"], c["W"], c["H"], c["pfx"], c["pfy"], want_samples=want_samples, **kw)


@pytest.mark.parametrize("name", render_cases())
def test_port_matches_reference_fixture(port, name):
    c = load_case(name)
    rgb, srgb, sprim = _render(port, c)
    assert np.array_equal(sprim, c["sample_prim"]), "primary primitive ids"
    assert np.array_equal(bits(srgb), bits(c["sample_rgb"])), "per-sample float RGB"
    assert np.array_equal(bits(rgb), bits(c["rgb"])), "per-pixel clamped float RGB"
    assert np.array_equal(port.quantise(rgb), c["u8"]), "u8 image (truncating quantiser)"


@pytest.mark.parametrize("name", ["cube_oblique_96_pf2", "glass_56_lvl6", "shadow_test_2lights_lvl3", "quirks_72_pf2"])
def test_reference_reproduces_fixture(ref, name):
    c = load_case(name)
    rgb, srgb, sprim = _render(ref, c)
    assert np.array_equal(sprim, c["sample_prim"])
    assert np.array_equal(bits(srgb), bits(c["sample_rgb"]))
    assert np.array_equal(bits(rgb), bits(c["rgb"]))


def test_trace_fixture(port):
    z = np.load(GOLDEN + "/trace_shadow_test.npz")
    port.set_scene(load_scene("shadow_test"))
    port.configure(z["eye"], [z["eye"]], 63, 10)
    rgb, prim, hit = port.trace(z["origins"], z["dests"])
    assert np.array_equal(prim, z["prim"])
    assert np.array_equal(bits(rgb), bits(z["rgb"]))
    h = prim >= 0
    assert np.array_equal(bits(hit[h]), bits(z["hit"][h]))


def test_threads_do_not_change_bits(port):
    c = load_case("room_64_pf2_lvl4")
    a = _render(port, c, threads=1)
    b = _render(port, c, threads=4)
    for x, y in zip(a, b):
        assert np.array_equal(bits(x) if x.dtype == np.float32 else x, bits(y) if y.dtype == np.float32 else y)


def test_row_subset_equals_full_rows(port):
    """y0/ystep (used by the CPU baseline's bounded sample and by the sharding tests) renders exactly those rows."""
    c = load_case("shadow_test_64_pf2")
    rgb, _, _ = _render(port, c, want_samples=False, y0=3, ystep=8)
    rows = np.arange(3, c["H"], 8)
    assert np.array_equal(bits(rgb[rows]), bits(c["rgb"][rows]))
    other = np.setdiff1d(np.arange(c["H"]), rows)
    assert not rgb[other].any()


def test_ray_counts(port):
    """intersectMesh calls by kind: one shadow ray per hit per light, one bounce per hit while lvl < max_lvl."""
    c = load_case("room_48_2lights_lvl10")
    port.reset_counts()
    _render(port, c, want_samples=False)
    primary, shadow, bounce = port.ray_counts()
    assert primary == c["W"] * c["H"] * c["pfx"] * c["pfy"]
    assert shadow % len(c["lights"]) == 0
    hits = shadow // len(c["lights"])
    assert hits >= np.count_nonzero(c["sample_prim"] >= 0)
    assert 0 < bounce <= hits


def test_spheres_extension_is_consistent(port):
    """No reference exists for analytic spheres: check the oracle's own definition against a
    finely tessellated sphere -- ids differ, geometry must agree to tessellation accuracy."""
    from raytracert_b200 import host, scenes
    s = scenes.mirror_room(n=12)
    sph = np.array([[0.0, 1.6, 0.3, 0.35, 2]], np.float32)
    cam = host.Camera(40, 40, (0.3, 1.6, 4.2), (0, 0.8, 0))
    port.set_scene(s)
    port.L.orc_set_spheres.argtypes = [__import__("ctypes").c_int, __import__("ctypes").c_void_p]
    port.L.orc_set_spheres(1, sph.ctypes.data)
    port.configure(cam.eye, [(1.5, 2.8, 2.5)], 63, 3)
    _, _, prim = port.render(cam.corners, 40, 40, 1, 1, want_samples=True)
    port.L.orc_set_spheres(0, sph.ctypes.data)
    n_sphere = np.count_nonzero(prim == s.n_triangles)
    # projected radius ~ 0.35 / 3.9 * (40 / (2 tan 25deg)) ~ 3.85 px -> ~46 px
    assert 30 <= n_sphere <= 64


def test_refraction_angle_threshold_matches_libm():
    """k_shade decides the reference's `acosf(check) <= 2` branch (raytracing.cpp:297-298) as `check >= T` with
    T = bits 0xbed51136, so that CUDA's acosf (which is not glibc's) can never flip reflect <-> refract.  Check T against
    this machine's libm (what the oracle and the reference build call): a window of 2e5 floats around T, a sweep of
    the whole range [-1, 0) in coarse steps, and the out-of-domain side."""
    import ctypes as C
    m = C.CDLL("libm.so.6")
    m.acosf.restype = C.c_float; m.acosf.argtypes = [C.c_float]
    T = 0xbed51136
    def f(bits):
        return np.array([bits], np.uint32).view(np.float32)[0]
    for b in list(range(T - 100000, T + 100000, 7)) + list(range(T - 64, T + 64)) + list(range(0x80000001, 0xbf800000, 1 << 14)):
        x = f(b)
        assert (m.acosf(x) <= 2.0) == (x >= f(T)), hex(b)
    assert not (m.acosf(np.float32(-1.0000001)) <= 2.0)      # NaN: refraction branch on both sides


def test_headline_pin_is_the_bench_frame(port):
    """tests/golden/pins/balls.npz (the whole 800x800x16 headline frame rendered by the unmodified reference,
    tools/make_headline_pin.py) must describe the frame bench.py renders today: same scene bytes, camera, lights -- and the
    port oracle, which the fixtures pin to the reference bit for bit, must reproduce one of its rows (ids and u8)."""
    import os, sys, zlib
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    import bench
    from raytracert_b200 import host
    path = os.path.join(ROOT, "tests", "golden", "pins", "balls.npz")
    if not os.path.exists(path):
        pytest.skip("pin not generated")
    z = np.load(path)
    scene, W, H, pf, lvl, eye, center, lights, _ = bench.workload("balls")
    cam = host.Camera(W, H, eye, center)
    c = 0
    for a in (scene.vertices, scene.indices, scene.tri_material, scene.materials):
        c = zlib.crc32(np.ascontiguousarray(a).tobytes(), c)
    assert np.uint32(c) == z["scene_crc"] and int(z["n_triangles"]) == scene.n_triangles
    assert np.array_equal(cam.corners, z["corners"]) and np.array_equal(np.asarray(lights, np.float32), z["lights"])
    assert (W, H, pf, lvl) == (int(z["W"]), int(z["H"]), int(z["pf"]), int(z["max_lvl"])) and len(z["rows"]) == H
    y = 470
    port.set_scene(scene); port.configure(cam.eye, lights, 63, lvl)
    rgb, _, prim = port.render(cam.corners, W, H, pf, pf, y0=y, ystep=H, want_samples=True)
    row_ids = prim.reshape(H, -1)[y]
    assert np.uint32(zlib.crc32(row_ids.astype("<i4").tobytes())) == z["id_crc"][y]
    ids = np.frombuffer(zlib.decompress(z["ids_z"].tobytes()), "<i4").reshape(H, -1)
    assert np.array_equal(ids[y], row_ids)
    assert np.array_equal(port.quantise(rgb[y]), z["u8"][y])
