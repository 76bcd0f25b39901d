"""Deterministic synthetic scenes (no RNG) for the configurations of BASELINE.json.

The headline scene ``Balls.obj`` is NOT in the reference checkout (/root/reference/.MISSING_LARGE_BLOBS);
``balls_standin`` is a documented stand-in: an "island" height field plus three tessellated spheres, using
the four materials of the reference's Balls.mtl (values restated in BALLS_MATERIALS).  Every report that
uses it says "stand-in".

All generators return :class:`raytracert_b200.host.Scene` (float32 vertices, uint32 indices).  ``write_obj``
emits OBJ+MTL text so the same scene can go through the C++ loader (host/mesh.cpp).
"""
import os

import numpy as np

from .host import Scene, face_normals

ALL_FLAGS = 63.0
# Kd Ns | Ka Ni | Ks Tr | flags : values of the reference's CG_Project/Balls.mtl:4-38
BALLS_MATERIALS = {
    "Material.002": [0.002, 1.0, 0.0, 96.078431, 0, 0, 0, 1.0, 0.5, 0.5, 0.5, 1.0, ALL_FLAGS, 0, 0, 0],
    "Material.003": [0.267942, 0.273673, 0.281009, 96.078431, 0, 0, 0, 1.0, 0.5, 0.5, 0.5, 1.0, ALL_FLAGS, 0, 0, 0],
    "Material.004": [0.420025, 0.420025, 0.420025, 96.078431, 0, 0, 0, 1.0, 0.5, 0.5, 0.5, 1.0, ALL_FLAGS, 0, 0, 0],
    "Material.005": [0.110282, 0.273831, 0.067133, 96.078431, 0, 0, 0, 1.0, 0.5, 0.5, 0.5, 1.0, ALL_FLAGS, 0, 0, 0],
}
# the loader's built-in material #0 (mesh.cpp:107-117): Kd .5, Ka 0, Ks .5, Ns 96.7, no Ni / Tr (pinned 1)
DEFAULT_MATERIAL = [0.5, 0.5, 0.5, 96.7, 0, 0, 0, 1.0, 0.5, 0.5, 0.5, 1.0, 15.0, 0, 0, 0]


def _finish(vertices, indices, tri_material, materials, names):
    vertices = np.asarray(vertices, np.float32)
    indices = np.asarray(indices, np.uint32)
    return Scene(vertices, indices, np.asarray(tri_material, np.uint32), face_normals(vertices, indices),
                 np.asarray(materials, np.float32), list(names))


def uv_sphere(center, radius, slices, stacks):
    """Closed UV sphere: slices*(stacks-2)*2 + 2*slices triangles, outward winding."""
    cx, cy, cz = center
    verts = [(cx, cy + radius, cz)]
    for i in range(1, stacks):
        th = np.pi * i / stacks
        y, r = np.cos(th), np.sin(th)
        ph = 2.0 * np.pi * np.arange(slices) / slices
        ring = np.stack([cx + radius * r * np.cos(ph), np.full(slices, cy + radius * y), cz + radius * r * np.sin(ph)], 1)
        verts.extend(map(tuple, ring))
    verts.append((cx, cy - radius, cz))
    verts = np.asarray(verts, np.float64)
    j = np.arange(slices)
    jn = (j + 1) % slices
    tris = [np.stack([np.zeros(slices, np.int64), 1 + jn, 1 + j], 1)]
    for i in range(stacks - 2):
        a = 1 + i * slices
        b = a + slices
        tris.append(np.stack([a + j, a + jn, b + j], 1))
        tris.append(np.stack([a + jn, b + jn, b + j], 1))
    last = 1 + (stacks - 1) * slices
    a = 1 + (stacks - 2) * slices
    tris.append(np.stack([np.full(slices, last), a + j, a + jn], 1))
    return verts, np.concatenate(tris)


def balls_standin(grid=128, slices=64, stacks=32):
    """Island height field (grid x grid cells -> 2*grid^2 triangles) + three UV spheres.
    Default: 32768 + 3*3968 = 44672 triangles."""
    u = np.linspace(-3.0, 3.0, grid + 1)
    X, Z = np.meshgrid(u, u, indexing="xy")
    r2 = X * X + Z * Z
    Y = 0.95 * np.exp(-r2 / (1.7 ** 2)) * (1.0 + 0.22 * np.sin(2.7 * X + 0.4) * np.cos(2.3 * Z - 0.3)) \
        + 0.03 * np.sin(6.1 * X) * np.sin(5.3 * Z) - 0.12
    Y = np.maximum(Y, 0.0)  # flat "sea" around the island
    verts = [np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1)]
    ii, jj = np.meshgrid(np.arange(grid), np.arange(grid), indexing="xy")
    a = (jj * (grid + 1) + ii).ravel()
    b = a + 1
    c = a + grid + 1
    d = c + 1
    tris = [np.concatenate([np.stack([a, c, b], 1), np.stack([b, c, d], 1)])]
    mats = [np.full(2 * grid * grid, 4, np.uint32)]  # Material.005 (dark green terrain)
    base = (grid + 1) ** 2
    for k, (center, rad) in enumerate([((-0.95, 1.25, 0.35), 0.42), ((0.15, 1.45, -0.55), 0.42), ((1.05, 1.15, 0.55), 0.42)]):
        v, t = uv_sphere(center, rad, slices, stacks)
        verts.append(v)
        tris.append(t + base)
        mats.append(np.full(len(t), 1 + k, np.uint32))  # Material.002 / .003 / .004
        base += len(v)
    names = ["StandardMaterialInitFromTriMesh"] + list(BALLS_MATERIALS)
    materials = [DEFAULT_MATERIAL] + [BALLS_MATERIALS[n] for n in BALLS_MATERIALS]
    return _finish(np.concatenate(verts), np.concatenate(tris), np.concatenate(mats), materials, names)


def balls_with_sphere_primitives(grid=128):
    """The Balls stand-in with its three balls as analytic `Sphere` primitives (Sphere.h of the reference; own
    semantics) instead of tessellated meshes: island height field (2*grid^2 triangles) + 3 spheres."""
    full = balls_standin(grid=grid, slices=8, stacks=4)
    nt = 2 * grid * grid                       # the terrain comes first in balls_standin
    nv = (grid + 1) ** 2
    s = Scene(full.vertices[:nv], full.indices[:nt], full.tri_material[:nt], full.normals[:nt], full.materials, full.names)
    s.spheres = np.array([[-0.95, 1.25, 0.35, 0.42, 1], [0.15, 1.45, -0.55, 0.42, 2], [1.05, 1.15, 0.55, 0.42, 3]], np.float32)
    return s


def tessellated_sphere(slices=1000, stacks=501, ground=False):
    """BASELINE C4: UV sphere of radius 1 at the origin; slices=1000, stacks=501 -> exactly 1,000,000 triangles."""
    v, t = uv_sphere((0.0, 0.0, 0.0), 1.0, slices, stacks)
    mats = np.full(len(t), 1, np.uint32)
    if ground:
        g = np.array([[-4, -1.2, -4], [4, -1.2, -4], [-4, -1.2, 4], [4, -1.2, 4]], np.float64)
        t = np.concatenate([t, np.array([[0, 2, 1], [1, 2, 3]]) + len(v)])
        v = np.concatenate([v, g])
        mats = np.concatenate([mats, np.full(2, 2, np.uint32)])
    names = ["StandardMaterialInitFromTriMesh", "Material.002", "Material.003"]
    materials = [DEFAULT_MATERIAL, BALLS_MATERIALS["Material.002"], BALLS_MATERIALS["Material.003"]]
    return _finish(v, t, mats, materials, names)


def unit_cube():
    """The 12-triangle unit cube with four coloured materials, same topology as the reference's cube.obj
    (vertex (x,y,z) in {0,1}^3, index = 4x+2y+z) with cube.mtl's colours; Tr/Ni pinned to 1."""
    v = np.array([[x, y, z] for x in (0, 1) for y in (0, 1) for z in (0, 1)], np.float64)
    f = np.array([[1, 7, 5], [1, 3, 7], [1, 4, 3], [1, 2, 4], [3, 8, 7], [3, 4, 8], [5, 7, 8], [5, 8, 6],
                  [1, 5, 6], [1, 6, 2], [2, 6, 8], [2, 8, 4]]) - 1
    m = np.array([1, 1, 4, 4, 2, 2, 3, 3, 2, 2, 1, 1], np.uint32)

    def mk(kd):
        return [kd[0], kd[1], kd[2], 5.0, 0, 0, 0, 1.0, kd[0], kd[1], kd[2], 1.0, 15.0, 0, 0, 0]
    materials = [DEFAULT_MATERIAL, mk((.5, .5, .5)), mk((.8, 0, 0)), mk((0, .8, 0)), mk((0, 0, .8))]
    names = ["StandardMaterialInitFromTriMesh", "buffy-gray", "buffy-red", "buffy-green", "buffy-blue"]
    return _finish(v, f, m, materials, names)


def mirror_room(n=24):
    """Small closed test scene: a floor, a back wall with a mirror-like Ks, and a torus-ish ring of
    spheres' worth of tessellated balls; exercises shadows + multi-bounce reflection cheaply."""
    verts, tris, mats = [], [], []
    base = 0

    def quad(p0, p1, p2, p3, m, div):
        nonlocal base
        s = np.linspace(0, 1, div + 1)
        S, T = np.meshgrid(s, s, indexing="xy")
        P = (np.asarray(p0)[None] * ((1 - S) * (1 - T)).ravel()[:, None] + np.asarray(p1)[None] * (S * (1 - T)).ravel()[:, None]
             + np.asarray(p2)[None] * ((1 - S) * T).ravel()[:, None] + np.asarray(p3)[None] * (S * T).ravel()[:, None])
        ii, jj = np.meshgrid(np.arange(div), np.arange(div), indexing="xy")
        a = (jj * (div + 1) + ii).ravel() + base
        verts.append(P)
        tris.append(np.concatenate([np.stack([a, a + 1, a + div + 1], 1), np.stack([a + 1, a + div + 2, a + div + 1], 1)]))
        mats.append(np.full(2 * div * div, m, np.uint32))
        base += len(P)

    quad((-2, 0, 2), (2, 0, 2), (-2, 0, -2), (2, 0, -2), 1, 6)          # floor (normal +y)
    quad((-2, 0, -2), (2, 0, -2), (-2, 3, -2), (2, 3, -2), 2, 4)        # back wall (normal +z), mirror
    quad((-2, 0, 2), (-2, 0, -2), (-2, 3, 2), (-2, 3, -2), 3, 4)        # left wall (normal +x)
    for k, c in enumerate([(-0.7, 0.5, 0.2), (0.6, 0.45, -0.4), (0.1, 0.9, 0.7)]):
        v, t = uv_sphere(c, 0.45 - 0.05 * k, n, n // 2)
        verts.append(v)
        tris.append(t + base)
        mats.append(np.full(len(t), 1 + (k + 1) % 3, np.uint32))
        base += len(v)
    materials = [DEFAULT_MATERIAL,
                 [0.26, 0.64, 0.13, 96.078431, 0.02, 0.02, 0.02, 1.0, 0.47, 0.48, 0.5, 1.0, ALL_FLAGS, 0, 0, 0],
                 [0.64, 0.64, 0.64, 30.0, 0, 0, 0, 1.0, 0.9, 0.9, 0.9, 1.0, ALL_FLAGS, 0, 0, 0],
                 [0.06, 0.16, 0.64, 96.078431, 0, 0, 0, 1.0, 0.5, 0.5, 0.5, 1.0, ALL_FLAGS, 0, 0, 0]]
    names = ["StandardMaterialInitFromTriMesh", "Floor", "Mirror", "Blue"]
    return _finish(np.concatenate(verts), np.concatenate(tris), np.concatenate(mats), materials, names)


def write_obj(scene, obj_path, mtl_name=None):
    """Emit OBJ + MTL text for `scene` (materials 1.. by name; material 0 is the loader's built-in)."""
    mtl_name = mtl_name or (os.path.splitext(os.path.basename(obj_path))[0] + ".mtl")
    d = os.path.dirname(os.path.abspath(obj_path))
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, mtl_name), "w") as f:
        f.write("# generated by raytracert_b200.scenes.write_obj\n\n")
        for i in range(1, len(scene.materials)):
            m = scene.materials[i]
            fl = int(m[12])
            f.write(f"newmtl {scene.names[i]}\n")
            if fl & 8: f.write(f"Ns {float(m[3])!r}\n")
            if fl & 2: f.write(f"Ka {float(m[4])!r} {float(m[5])!r} {float(m[6])!r}\n")
            if fl & 1: f.write(f"Kd {float(m[0])!r} {float(m[1])!r} {float(m[2])!r}\n")
            if fl & 4: f.write(f"Ks {float(m[8])!r} {float(m[9])!r} {float(m[10])!r}\n")
            if fl & 16: f.write(f"Ni {float(m[7])!r}\n")
            if fl & 32: f.write(f"d {float(m[11])!r}\n")
            f.write("illum 2\n\n")
    with open(obj_path, "w") as f:
        f.write(f"# generated by raytracert_b200.scenes.write_obj\nmtllib {mtl_name}\n")
        np.savetxt(f, scene.vertices.astype(np.float64), fmt="v %.9g %.9g %.9g")
        order = np.argsort(scene.tri_material, kind="stable")
        cur = None
        # keep triangle order: emit usemtl whenever the material changes
        for i in range(scene.n_triangles):
            m = int(scene.tri_material[i])
            if m != cur:
                if m == 0:
                    raise ValueError("write_obj: triangles on the built-in material 0 cannot be named")
                f.write(f"usemtl {scene.names[m]}\n")
                cur = m
            a, b, c = (int(x) + 1 for x in scene.indices[i])
            f.write(f"f {a} {b} {c}\n")
        del order
    return obj_path
