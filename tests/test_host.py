"""Host side (C++ in the reference's style, raytracert_b200/host): OBJ/MTL loader against dumps of the
reference's own loader, face normals, camera corner rays, PPM writer."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, bits, load_scene

REF_DIR = "/root/reference/CG_Project"


def _check_against_dump(s, z):
    assert np.array_equal(bits(s.vertices), bits(z["vertices"]))
    assert np.array_equal(s.indices, z["indices"])
    assert np.array_equal(s.tri_material, z["tri_material"])
    assert np.array_equal(bits(s.normals), bits(z["normals"]))
    assert list(s.names) == [str(n) for n in z["names"]]
    assert np.array_equal(s.materials[:, 12], z["flags"])
    flags = z["flags"].astype(int)
    for i, fl in enumerate(flags):
        # every flagged field must match; unflagged scalars are pinned (Tr = Ni = 1)
        for bit, cols in [(1, [0, 1, 2]), (2, [4, 5, 6]), (4, [8, 9, 10]), (8, [3]), (16, [7]), (32, [11])]:
            if fl & bit:
                assert np.array_equal(bits(s.materials[i, cols]), bits(z["materials"][i, cols])), (i, bit)
        if not fl & 32: assert s.materials[i, 11] == 1.0
        if not fl & 16: assert s.materials[i, 7] == 1.0
    # leak semantics (mesh.h:43-53): unflagged colours keep the previous material's value -- they matter
    # because shade() multiplies by Ks without looking at has_Ks (raytracing.cpp:363)
    assert np.array_equal(bits(s.materials[:, 8:11]), bits(z["materials"][:, 8:11]))


def test_quirks_obj_matches_reference_loader(built):
    from raytracert_b200 import host
    s = host.load_obj(os.path.join(GOLDEN, "obj", "quirks.obj"))
    z = np.load(os.path.join(GOLDEN, "loader", "quirks.npz"))
    assert s.n_triangles == 15 and len(s.vertices) == 14   # quad x4 -> 8, tris x4, pentagon -> 3, "f 1 2" dropped
    _check_against_dump(s, z)


@pytest.mark.skipif(not os.path.exists(REF_DIR), reason="reference assets are only in the build container")
@pytest.mark.parametrize("obj,fixture", [("cube.obj", "cube"), ("dodgeColorTest.obj", "dodge"), ("Models/shadow_test.obj", "shadow_test")])
def test_shipped_scenes_match_fixtures(built, obj, fixture):
    from raytracert_b200 import host
    s = host.load_obj(os.path.join(REF_DIR, obj))
    g = load_scene(fixture)
    assert np.array_equal(bits(s.vertices), bits(g.vertices)) and np.array_equal(s.indices, g.indices)
    assert np.array_equal(s.tri_material, g.tri_material) and np.array_equal(bits(s.normals), bits(g.normals))
    assert np.array_equal(bits(s.materials[:, :13]), bits(g.materials[:, :13]))


def test_loader_edge_cases(built, tmp_path):
    from raytracert_b200 import host
    with pytest.raises(FileNotFoundError):
        host.load_obj(str(tmp_path / "missing.obj"))       # the reference crashes in fclose(NULL); pinned to an error
    # no mtllib, face before any usemtl, unknown usemtl -> material 0 (pinned UB), CRLF line ends, long line split
    p = tmp_path / "edge.obj"
    long_comment = "#" + "x" * 300 + "\n"                   # > 255 chars: the tail becomes a "line" starting with 'x' -> ignored
    p.write_bytes(("v 0 0 0\r\nv 1 0 0\r\nv 0 1 0\r\nv 1 1 0\r\n" + long_comment +
                   "f 1 2 3\r\nusemtl nothing\r\nf 2/1 4/1 3/1\r\nf 1 2\r\n").encode())
    s = host.load_obj(str(p))
    assert s.n_triangles == 2 and list(s.tri_material) == [0, 0]
    assert np.array_equal(s.indices, [[0, 1, 2], [1, 3, 2]])
    assert len(s.materials) == 1 and s.materials[0, 12] == 15 and abs(s.materials[0, 3] - 96.7) < 1e-5
    # out-of-range vertex index: the reference reads out of bounds; here the load is refused
    q = tmp_path / "oob.obj"
    q.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 9\n")
    with pytest.raises(FileNotFoundError):
        host.load_obj(str(q))


def test_write_obj_roundtrip(built, tmp_path):
    from raytracert_b200 import host, scenes
    s = scenes.mirror_room(n=8)
    path = scenes.write_obj(s, str(tmp_path / "room.obj"))
    t = host.load_obj(path)
    assert np.array_equal(bits(t.vertices), bits(s.vertices)) and np.array_equal(t.indices, s.indices)
    assert np.array_equal(t.tri_material, s.tri_material) and np.array_equal(bits(t.normals), bits(s.normals))
    assert np.allclose(t.materials[1:, :12], s.materials[1:, :12], rtol=1e-6)


def test_face_normals_match_oracle(built, port):
    from raytracert_b200 import host
    s = load_scene("dodge")
    port.set_scene(s)
    out = np.zeros((s.n_triangles, 3), np.float32)
    port.L.orc_get_normals.argtypes = [__import__("ctypes").c_void_p]
    port.L.orc_get_normals(out.ctypes.data)
    assert np.array_equal(bits(host.face_normals(s.vertices, s.indices)), bits(out))
    assert np.array_equal(bits(s.normals), bits(out))


def test_default_camera(built):
    """modelview T(0,0,-4), gluPerspective(50, W/H, 1, 10) (main.cpp:217-219, 294): eye (0,0,4); corner origins lie on
    the near plane z = 3, dests on the far plane z = -6; corner (0,0) is the top-left of the image."""
    from raytracert_b200 import host
    cam = host.Camera(800, 800)
    assert np.allclose(cam.eye, [0, 0, 4])
    c = cam.corners.reshape(4, 2, 3)
    assert np.allclose(c[:, 0, 2], 3.0, atol=1e-5) and np.allclose(c[:, 1, 2], -6.0, atol=1e-4)
    t = np.tan(np.radians(25.0))
    assert np.allclose(c[0, 0, :2], [-t, t], atol=1e-5)                      # window (0, H): x = -1, y = +1 in NDC
    assert np.allclose(c[3, 0, :2], [t * (2 * 799 / 800 - 1), -t * (1 - 2 / 800)], atol=1e-5)  # asymmetric by design
    assert np.allclose(c[:, 1, :2], 10 * c[:, 0, :2], rtol=1e-5)
    cam2 = host.Camera(1920, 1080, (.75, .55, 1.1), (.07, 0, .23))
    assert np.allclose(cam2.eye, [.75, .55, 1.1], atol=1e-6)
    d = np.linalg.norm(cam2.corners.reshape(4, 2, 3)[:, 0] - cam2.eye, axis=1)
    assert np.all(d > 1.0) and np.all(d < 1.6)                               # near plane is 1 away along the axis


def test_ppm_writer(built, tmp_path):
    """P6 header + truncating quantiser (main.cpp:112-117): 1.0 -> 255, 0.999 -> 254, 0.5 -> 127."""
    from raytracert_b200 import host
    rgb = np.zeros((2, 3, 3), np.float32)
    rgb[0, 0] = [1.0, 0.999, 0.5]
    rgb[1, 2] = [0.0039, 0.00393, 0.25]
    p = tmp_path / "a.ppm"
    host.write_ppm(str(p), rgb, 3, 2)
    raw = p.read_bytes()
    head = b"P6\n3 2\n255\n"
    assert raw.startswith(head) and len(raw) == len(head) + 18
    px = np.frombuffer(raw[len(head):], np.uint8).reshape(2, 3, 3)
    assert list(px[0, 0]) == [255, 254, 127] and list(px[1, 2]) == [0, 1, 63]


def test_cpp_app_fails_loudly_without_gpu(built):
    """The C++ drop-in has no CPU path: without a device it must say so and exit non-zero."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    app = os.path.join(ROOT, "raytracert_b200", "_build", "rt_main")
    if not os.path.exists(app):
        subprocess.run(["make", "-C", ROOT, "app"], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run([app, os.path.join(GOLDEN, "obj", "quirks.obj"), "--size", "16x16"], capture_output=True, text=True, cwd=str(ROOT))
    assert r.returncode != 0 and "no CPU fallback" in r.stdout
