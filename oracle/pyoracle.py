"""ctypes drivers for the two oracles -- TEST INFRASTRUCTURE ONLY.

  * ``RefOracle``  : oracle/_ref/libref_oracle.so  -- the UNMODIFIED reference sources behind a headless shim
  * ``PortOracle`` : oracle/_build/librt_oracle.so -- the plain-C restatement (oracle/rt_oracle.c)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Both classes expose the same calls so a test can swap one for the other.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libref_oracle.so")
PORT_SO = os.path.join(HERE, "_build", "librt_oracle.so")

FEATURES_ALL = 63  # ambient 1 | diffuse 2 | specular 4 | reflection 8 | shadows 16 | refraction 32


class _Oracle:
    prefix = None
    so = None

    def __init__(self):
        if not os.path.exists(self.so):
            raise FileNotFoundError(self.so)
        self.L = C.CDLL(self.so)
        self.kind = self.prefix
        self.n_triangles = 0

    def _f(self, name):
        return getattr(self.L, f"{self.prefix}_{name}")

    def set_scene(self, scene):
        v = np.ascontiguousarray(scene.vertices, np.float32)
        idx = np.ascontiguousarray(scene.indices, np.uint32)
        mat = np.ascontiguousarray(scene.tri_material, np.uint32)
        mats = np.ascontiguousarray(scene.materials, np.float32)
        f = self._f("set_scene")
        f.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        f(len(v), v.ctypes.data, len(idx), idx.ctypes.data, mat.ctypes.data, len(mats), mats.ctypes.data)
        self.n_triangles = len(idx)

    def configure(self, eye, lights, features=FEATURES_ALL, max_lvl=10):
        eye = np.ascontiguousarray(eye, np.float32)
        lights = np.ascontiguousarray(lights, np.float32).reshape(-1, 3)
        f = self._f("set_camera"); f.argtypes = [C.c_void_p]; f(eye.ctypes.data)
        f = self._f("set_lights"); f.argtypes = [C.c_int, C.c_void_p]; f(len(lights), lights.ctypes.data)
        f = self._f("set_toggles"); f.argtypes = [C.c_int] * 6
        f(*(1 if features & b else 0 for b in (1, 2, 4, 8, 16, 32)))
        f = self._f("set_max_lvl"); f.argtypes = [C.c_int]; f(int(max_lvl))

    def render(self, corners, W, H, pfx=1, pfy=1, y0=0, ystep=1, want_samples=False, threads=0, x0=0, xstep=1):
        """Renders the pixel lattice rows y0::ystep x columns x0::xstep (default: the full frame).
        Returns (rgb[H,W,3] clamped float32, sample_rgb or None, sample_prim or None)."""
        corners = np.ascontiguousarray(corners, np.float32)
        rgb = np.zeros((H, W, 3), np.float32)
        srgb = np.zeros((H * W * pfx * pfy, 3), np.float32) if want_samples else None
        sprim = np.full(H * W * pfx * pfy, -2, np.int32) if want_samples else None
        f = self._f("render")
        f.argtypes = [C.c_void_p] + [C.c_int] * 8 + [C.c_void_p] * 3 + [C.c_int]
        f(corners.ctypes.data, W, H, pfx, pfy, y0, ystep, x0, xstep, rgb.ctypes.data,
          srgb.ctypes.data if want_samples else None, sprim.ctypes.data if want_samples else None,
          threads or (os.cpu_count() or 1))
        return rgb, srgb, sprim

    def trace(self, origins, dests):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dests, np.float32).reshape(-1, 3)
        n = len(o)
        rgb = np.zeros((n, 3), np.float32)
        prim = np.zeros(n, np.int32)
        hit = np.zeros((n, 3), np.float32)
        f = self._f("trace")
        f.argtypes = [C.c_int] + [C.c_void_p] * 5
        f(n, o.ctypes.data, d.ctypes.data, rgb.ctypes.data, prim.ctypes.data, hit.ctypes.data)
        return rgb, prim, hit

    def quantise(self, rgb):
        rgb = np.ascontiguousarray(rgb, np.float32)
        out = np.zeros(rgb.size, np.uint8)
        f = self._f("quantise"); f.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        f(rgb.ctypes.data, rgb.size, out.ctypes.data)
        return out.reshape(rgb.shape)


class RefOracle(_Oracle):
    """The real reference (raytracing.cpp + mesh.cpp) behind oracle/ref_harness.cpp."""
    prefix = "ref"
    so = REF_SO

    def load_obj(self, path):
        """Runs the reference's own init()/loadMesh on an OBJ and returns the dumped arrays."""
        f = self.L.ref_load_obj; f.argtypes = [C.c_char_p]
        nt = f(os.fsencode(path))
        if nt < 0:
            raise FileNotFoundError(path)
        nv, nt, nm = C.c_int(), C.c_int(), C.c_int()
        self.L.ref_counts(C.byref(nv), C.byref(nt), C.byref(nm))
        v = np.zeros((nv.value, 3), np.float32); idx = np.zeros((nt.value, 3), np.uint32)
        mat = np.zeros(nt.value, np.uint32); nrm = np.zeros((nt.value, 3), np.float32)
        self.L.ref_get_vertices.argtypes = [C.c_void_p]; self.L.ref_get_vertices(v.ctypes.data)
        self.L.ref_get_triangles.argtypes = [C.c_void_p, C.c_void_p]; self.L.ref_get_triangles(idx.ctypes.data, mat.ctypes.data)
        self.L.ref_get_normals.argtypes = [C.c_void_p]; self.L.ref_get_normals(nrm.ctypes.data)
        mats = np.zeros((nm.value, 16), np.float32); names = []
        self.L.ref_get_material.argtypes = [C.c_int, C.c_void_p, C.c_char_p, C.c_int]
        for i in range(nm.value):
            buf = C.create_string_buffer(256)
            self.L.ref_get_material(i, mats[i].ctypes.data, buf, 256)
            names.append(buf.value.decode("latin1"))
        self.n_triangles = nt.value
        return dict(vertices=v, indices=idx, tri_material=mat, normals=nrm, materials=mats, names=names)

    def pin_material(self, i, tr=None, ni=None):
        if tr is not None:
            self.L.ref_set_material_tr.argtypes = [C.c_int, C.c_float]; self.L.ref_set_material_tr(i, tr)
        if ni is not None:
            self.L.ref_set_material_ni.argtypes = [C.c_int, C.c_float]; self.L.ref_set_material_ni(i, ni)


class PortOracle(_Oracle):
    """Plain-C restatement of the reference algorithm (oracle/rt_oracle.c)."""
    prefix = "orc"
    so = PORT_SO

    def ray_counts(self):
        """(primary, shadow, bounce) intersectMesh-equivalent calls since the last reset."""
        out = (C.c_uint64 * 3)()
        self.L.orc_get_counts(out)
        return tuple(int(x) for x in out)

    def reset_counts(self):
        self.L.orc_reset_counts()


def available():
    return {"ref": os.path.exists(REF_SO), "port": os.path.exists(PORT_SO)}
