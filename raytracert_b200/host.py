"""ctypes view of ``librt_host.so`` (C++ host side: OBJ/MTL loader, flatten, camera, PPM writer).

Mirrors what the reference's skeleton does around the hot path: ``init()`` loads the mesh
(raytracing.cpp:42-73), ``produceRay`` makes the corner rays (main.cpp:300-320, 355-358) and
``Image::writeImage`` writes the PPM (main.cpp:102-128).
"""
import ctypes as C
import os

import numpy as np

from . import BUILD_DIR

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(BUILD_DIR, "librt_host.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `make host` (or __graft_entry__.build())")
        L = C.CDLL(path)
        L.rth_load_obj.restype = C.c_void_p
        L.rth_load_obj.argtypes = [C.c_char_p]
        L.rth_free.argtypes = [C.c_void_p]
        L.rth_counts.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 3
        L.rth_get_vertices.argtypes = [C.c_void_p, C.c_void_p]
        L.rth_get_triangles.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.rth_get_normals.argtypes = [C.c_void_p, C.c_void_p]
        L.rth_get_material.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_char_p, C.c_int]
        L.rth_face_normals.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.rth_scene.restype = C.c_void_p
        L.rth_scene.argtypes = [C.c_void_p]
        L.rth_corner_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.rth_default_camera.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.rth_lookat_camera.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_int] + [C.c_void_p] * 3
        L.rth_write_ppm.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        _LIB = L
    return _LIB


class Scene:
    """Flat scene in numpy form: what the loader produced, ready for rt_upload_scene / the oracles."""

    def __init__(self, vertices, indices, tri_material, normals, materials, names=None, spheres=None):
        self.vertices = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
        self.indices = np.ascontiguousarray(indices, np.uint32).reshape(-1, 3)
        self.tri_material = np.ascontiguousarray(tri_material, np.uint32).reshape(-1)
        self.normals = np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
        self.materials = np.ascontiguousarray(materials, np.float32).reshape(-1, 16)
        self.names = names or [f"m{i}" for i in range(len(self.materials))]
        # spheres: (n, 5) float32 rows cx cy cz radius material
        self.spheres = np.zeros((0, 5), np.float32) if spheres is None else np.ascontiguousarray(spheres, np.float32).reshape(-1, 5)

    @property
    def n_triangles(self):
        return len(self.indices)

    def corner(self, k):
        """float4-padded corner array k (0..2) of every triangle."""
        out = np.zeros((self.n_triangles, 4), np.float32)
        out[:, :3] = self.vertices[self.indices[:, k]]
        return out

    def flat(self):
        """The scene as the SoA float4 host buffers of ``rt_scene`` (what host/flatten.h produces once per load in the C++
        drop-in): corners v0/v1/v2, normal, material index.  Built once and kept -- these ARE the host buffers every
        rt_upload_scene call reads."""
        f = getattr(self, "_flat", None)
        if f is None:
            n = self.n_triangles
            nrm = np.zeros((n, 4), np.float32)
            nrm[:, :3] = self.normals
            f = self._flat = dict(v0=self.corner(0), v1=self.corner(1), v2=self.corner(2), normal=nrm,
                                  tri_material=np.ascontiguousarray(self.tri_material, np.uint32))
        return f

    def save(self, path):
        np.savez_compressed(path, vertices=self.vertices, indices=self.indices, tri_material=self.tri_material,
                            normals=self.normals, materials=self.materials, names=np.array(self.names),
                            spheres=self.spheres)

    @staticmethod
    def load(path):
        z = np.load(path, allow_pickle=False)
        return Scene(z["vertices"], z["indices"], z["tri_material"], z["normals"], z["materials"],
                     [str(s) for s in z["names"]], z["spheres"] if "spheres" in z.files else None)


def face_normals(vertices, indices):
    """Unit face normals with the reference's arithmetic (host/flatten.h append_face_normals)."""
    v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
    idx = np.ascontiguousarray(indices, np.uint32).reshape(-1, 3)
    out = np.zeros((len(idx), 3), np.float32)
    lib().rth_face_normals(len(v), v.ctypes.data, len(idx), idx.ctypes.data, out.ctypes.data)
    return out


def load_obj(path):
    """OBJ/MTL -> Scene through the C++ loader (host/mesh.cpp) + face normals (host/flatten.h)."""
    L = lib()
    h = L.rth_load_obj(os.fsencode(path))
    if not h:
        raise FileNotFoundError(f"cannot load OBJ {path}")
    try:
        nv, nt, nm = C.c_int(), C.c_int(), C.c_int()
        L.rth_counts(h, C.byref(nv), C.byref(nt), C.byref(nm))
        v = np.zeros((nv.value, 3), np.float32)
        idx = np.zeros((nt.value, 3), np.uint32)
        mat = np.zeros(nt.value, np.uint32)
        nrm = np.zeros((nt.value, 3), np.float32)
        L.rth_get_vertices(h, v.ctypes.data)
        L.rth_get_triangles(h, idx.ctypes.data, mat.ctypes.data)
        L.rth_get_normals(h, nrm.ctypes.data)
        mats = np.zeros((nm.value, 16), np.float32)
        names = []
        for i in range(nm.value):
            buf = C.create_string_buffer(256)
            L.rth_get_material(h, i, mats[i].ctypes.data, buf, 256)
            names.append(buf.value.decode("latin1"))
        return Scene(v, idx, mat, nrm, mats, names)
    finally:
        L.rth_free(h)


class Camera:
    """Modelview/projection pair + the 24 corner floats and the eye, for a W x H frame."""

    def __init__(self, W, H, eye=None, center=None, up=(0.0, 1.0, 0.0)):
        L = lib()
        self.W, self.H = int(W), int(H)
        self.modelview = np.zeros(16, np.float64)
        self.projection = np.zeros(16, np.float64)
        self.eye = np.zeros(3, np.float32)
        if eye is None:
            L.rth_default_camera(self.W, self.H, self.modelview.ctypes.data, self.projection.ctypes.data, self.eye.ctypes.data)
        else:
            e = np.asarray(eye, np.float64)
            c = np.asarray(center, np.float64)
            u = np.asarray(up, np.float64)
            L.rth_lookat_camera(e.ctypes.data, c.ctypes.data, u.ctypes.data, self.W, self.H,
                                self.modelview.ctypes.data, self.projection.ctypes.data, self.eye.ctypes.data)
        self.corners = np.zeros(24, np.float32)
        if L.rth_corner_rays(self.modelview.ctypes.data, self.projection.ctypes.data, self.W, self.H, self.corners.ctypes.data) != 0:
            raise ValueError("singular camera")


def write_ppm(path, rgb, W, H):
    rgb = np.ascontiguousarray(rgb, np.float32)
    if lib().rth_write_ppm(os.fsencode(path), rgb.ctypes.data, W, H) != 0:
        raise OSError(f"cannot write {path}")
