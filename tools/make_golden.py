"""Generate the committed golden fixtures under tests/golden/ from the REAL reference (oracle/_ref:
the unmodified /root/reference/CG_Project/{raytracing,mesh}.cpp behind oracle/ref_harness.cpp).

Run in the build container only (needs /root/reference for the shipped OBJ scenes and a built
oracle/_ref):      make -C oracle ref && python tools/make_golden.py

What it writes (all small, all regenerated deterministically):
  tests/golden/scenes/<name>.npz    scenes as the reference's own loader produced them (cube, dodge,
                                    shadow_test, quirks) or as raytracert_b200.scenes generated them
                                    (room, glass, balls_small, balls_fine); UB pins applied (Tr/Ni = 1
                                    where the MTL never sets them)
  tests/golden/loader/<name>.npz    raw reference-loader dumps for the loader parity test
  tests/golden/renders/<case>.npz   reference outputs per render case: clamped float RGB per pixel, u8
                                    image, per-sample primary primitive id, per-sample float RGB
  tests/golden/trace_shadow_test.npz  performRayTracing on a seeded batch of free rays

Nothing here is used by the product path.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402
from raytracert_b200 import host, scenes  # noqa: E402

REF = "/root/reference/CG_Project"
G = os.path.join(ROOT, "tests", "golden")


def pin(mats):
    """scalars the MTL never set are indeterminate in the reference; pin Tr = Ni = 1, Ns = 0."""
    m = np.array(mats, np.float32)
    for row in m:
        fl = int(row[12])
        if not fl & 32: row[11] = 1.0
        if not fl & 16: row[7] = 1.0
        if not fl & 8: row[3] = 0.0 if not np.isfinite(row[3]) else row[3]
        row[13:] = 0.0
    return m


def ref_scene(R, path):
    d = R.load_obj(path)
    return d, host.Scene(d["vertices"], d["indices"], d["tri_material"], d["normals"], pin(d["materials"]), d["names"])


def glass_room():
    """mirror_room with one ball made of glass (d 0.4, Ni 1.5): exercises refraction + the
    transparent-occluder rule of isShadow (raytracing.cpp:253-256)."""
    s = scenes.mirror_room(n=16)
    mats = s.materials.copy()
    glass = np.array([[0.05, 0.05, 0.08, 60.0, 0, 0, 0, 1.5, 0.8, 0.8, 0.8, 0.4, 63, 0, 0, 0]], np.float32)
    mats = np.concatenate([mats, glass])
    tm = s.tri_material.copy()
    # the last ball (material 1 in mirror_room's cycle) becomes glass
    nball = 2 * 16 * (16 // 2 - 1)
    tm[-nball:] = len(mats) - 1
    return host.Scene(s.vertices, s.indices, tm, s.normals, mats, s.names + ["Glass"])


def main():
    R = pyoracle.RefOracle()
    for d in ("scenes", "loader", "renders"):
        os.makedirs(os.path.join(G, d), exist_ok=True)
    sc = {}
    for name, rel in [("cube", REF + "/cube.obj"), ("dodge", REF + "/dodgeColorTest.obj"),
                      ("shadow_test", REF + "/Models/shadow_test.obj"), ("quirks", os.path.join(G, "obj", "quirks.obj"))]:
        raw, sc[name] = ref_scene(R, rel)
        sc[name].save(os.path.join(G, "scenes", name + ".npz"))
        if name in ("cube", "quirks"):
            np.savez_compressed(os.path.join(G, "loader", name + ".npz"), vertices=raw["vertices"], indices=raw["indices"],
                                tri_material=raw["tri_material"], normals=raw["normals"], flags=raw["materials"][:, 12].copy(),
                                materials=pin(raw["materials"]), names=np.array(raw["names"]))
    sc["room"] = scenes.mirror_room()
    sc["glass"] = glass_room()
    sc["balls_small"] = scenes.balls_standin(grid=24, slices=16, stacks=8)
    sc["balls_fine"] = scenes.balls_standin(grid=48, slices=24, stacks=12)     # fine enough for the clause-free / pencil kernels
    for name in ("room", "glass", "balls_small", "balls_fine"):
        sc[name].save(os.path.join(G, "scenes", name + ".npz"))

    OBL = ((2.6, 2.4, 3.0), (.5, .5, .5))
    ST = ((1, 5, 7), (1, 1.2, .7))
    DG = ((.75, .55, 1.1), (.07, 0, .23))
    RM = ((0.3, 1.6, 4.2), (0, 0.8, 0))
    BL = ((0.0, 2.6, 5.2), (0.0, 0.55, 0.0))
    # name, scene, W, H, look-at or None, pfx, pfy, max_lvl, features, lights (None = one light at the eye)
    cases = [
        ("cube_default_64", "cube", 64, 64, None, 1, 1, 10, 63, None),
        ("cube_oblique_96_pf2", "cube", 96, 96, OBL, 2, 2, 10, 63, None),
        ("cube_nonsquare_pf3x2", "cube", 80, 40, OBL, 3, 2, 10, 63, None),
        ("cube_ambient_diffuse_only", "cube", 48, 48, OBL, 1, 1, 10, 1 | 2, None),
        ("cube_no_shadows", "cube", 48, 48, OBL, 1, 1, 10, 63 & ~16, None),
        ("cube_no_reflection", "cube", 48, 48, OBL, 1, 1, 10, 63 & ~8, None),
        ("cube_specular_only_lvl0", "cube", 48, 48, OBL, 1, 1, 0, 4 | 8, None),
        ("quirks_72_pf2", "quirks", 72, 72, ((3.4, 3.0, 4.6), (0.4, 0.2, 0.2)), 2, 2, 6, 63, [(3.0, 5.0, 4.0)]),
        ("shadow_test_64_pf2", "shadow_test", 64, 64, ST, 2, 2, 10, 63, None),
        ("shadow_test_2lights_lvl3", "shadow_test", 56, 56, ST, 1, 1, 3, 63, [(1, 5, 7), (-2.0, 4.0, 1.0)]),
        ("dodge_48x27", "dodge", 48, 27, DG, 1, 1, 10, 63, None),
        ("dodge_32x18_pf2_lvl2", "dodge", 32, 18, DG, 2, 2, 2, 63, [(0.9, 1.2, 1.4)]),
        ("room_64_pf2_lvl4", "room", 64, 64, RM, 2, 2, 4, 63, [(1.5, 2.8, 2.5)]),
        ("room_48_2lights_lvl10", "room", 48, 48, RM, 1, 1, 10, 63, [(1.5, 2.8, 2.5), (-1.0, 2.0, 1.0)]),
        ("glass_56_lvl6", "glass", 56, 56, RM, 1, 1, 6, 63, [(1.5, 2.8, 2.5)]),
        ("glass_40_pf2_norefraction", "glass", 40, 40, RM, 2, 2, 4, 63 & ~32, [(1.5, 2.8, 2.5)]),
        ("balls_small_64_pf2_lvl3", "balls_small", 64, 64, BL, 2, 2, 3, 63, [(2.5, 4.0, 3.0)]),
        ("balls_fine_96x64_pf2_2lights_lvl3", "balls_fine", 96, 64, BL, 2, 2, 3, 63, [(2.5, 4.0, 3.0), (0.0, 2.6, 5.2)]),
        ("balls_fine_default_camera_72_pf2", "balls_fine", 72, 72, None, 2, 2, 2, 63, None),
    ]
    for name, sname, W, H, look, pfx, pfy, lvl, feats, lights in cases:
        cam = host.Camera(W, H) if look is None else host.Camera(W, H, look[0], look[1])
        lights = np.asarray([cam.eye] if lights is None else lights, np.float32)
        R.set_scene(sc[sname])
        R.configure(cam.eye, lights, feats, lvl)
        rgb, srgb, sprim = R.render(cam.corners, W, H, pfx, pfy, want_samples=True)
        u8 = R.quantise(rgb)
        np.savez_compressed(os.path.join(G, "renders", name + ".npz"), scene=np.array(sname), corners=cam.corners, eye=cam.eye,
                            lights=lights, W=W, H=H, pfx=pfx, pfy=pfy, max_lvl=lvl, features=feats, rgb=rgb, u8=u8,
                            sample_prim=sprim, sample_rgb=srgb)
        print(f"{name:32s} hit {np.mean(sprim >= 0):.3f} mean rgb {rgb.mean():.4f}")

    # free rays through performRayTracing (raytracing.cpp:410): seeded, aimed roughly at the scene
    rng = np.random.default_rng(20141031)
    n = 600
    o = rng.uniform(-4, 6, (n, 3)).astype(np.float32)
    o[:, 1] = rng.uniform(0.5, 7, n)
    d = (rng.uniform(-1.5, 3.5, (n, 3)) * np.array([1, 0.6, 1])).astype(np.float32)
    R.set_scene(sc["shadow_test"])
    eye = np.array([1, 5, 7], np.float32)
    R.configure(eye, [eye], 63, 10)
    rgb, prim, hit = R.trace(o, d)
    np.savez_compressed(os.path.join(G, "trace_shadow_test.npz"), origins=o, dests=d, eye=eye, rgb=rgb, prim=prim, hit=hit)
    print("trace: hit fraction", np.mean(prim >= 0))


if __name__ == "__main__":
    main()
