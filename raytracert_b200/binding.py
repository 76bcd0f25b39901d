"""ctypes binding of ``librt_b200.so`` -- the C ABI of ``include/rt_b200.h``.

This is the call path a user of the reference's plugin interface takes (``init`` -> ``rt_upload_scene``,
the 'r' key -> ``rt_render`` + ``rt_download_framebuffer``, ``performRayTracing`` -> ``rt_trace``).
There is no fallback of any kind: a missing library or a missing GPU raises.
"""
import ctypes as C
import os

import numpy as np

from . import BUILD_DIR

RT_AMBIENT, RT_DIFFUSE, RT_SPECULAR, RT_REFLECTION, RT_SHADOWS, RT_REFRACTION = 1, 2, 4, 8, 16, 32
RT_ALL_FEATURES = 63
RT_MAX_LIGHTS = 16
RT_OPT_TILE_CULLING = 1
RT_OPT_PENCIL = 2
RT_OPT_PENCIL_ANY = 3
RT_OPT_GRAPH = 4
RT_OPT_PENCIL_REFLECT = 5
RT_OPT_SMALL_TRACE = 6
RT_OPT_PENCIL_THREAD = 7


class RtMaterial(C.Structure):
    _fields_ = [("Kd", C.c_float * 3), ("Ns", C.c_float), ("Ka", C.c_float * 3), ("Ni", C.c_float),
                ("Ks", C.c_float * 3), ("Tr", C.c_float), ("flags", C.c_uint32), ("pad", C.c_uint32 * 3)]


class RtSphere(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("radius", C.c_float), ("material", C.c_uint32), ("pad", C.c_uint32 * 3)]


class RtScene(C.Structure):
    _fields_ = [("n_triangles", C.c_uint32), ("v0", C.c_void_p), ("v1", C.c_void_p), ("v2", C.c_void_p),
                ("normal", C.c_void_p), ("tri_material", C.c_void_p), ("n_materials", C.c_uint32),
                ("materials", C.c_void_p), ("n_spheres", C.c_uint32), ("spheres", C.c_void_p)]


class RtParams(C.Structure):
    _fields_ = [("corners", C.c_float * 24), ("width", C.c_uint32), ("height", C.c_uint32),
                ("pixelfactor_x", C.c_uint32), ("pixelfactor_y", C.c_uint32), ("max_lvl", C.c_int32),
                ("features", C.c_uint32), ("camera", C.c_float * 3), ("n_lights", C.c_uint32),
                ("lights", (C.c_float * 3) * RT_MAX_LIGHTS), ("want_prim_id", C.c_uint32)]


class RtStats(C.Structure):
    _fields_ = [("primary_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("bounce_rays", C.c_uint64),
                ("tri_tests", C.c_uint64), ("exact_evals", C.c_uint64), ("ms_total", C.c_float),
                ("ms_trace", C.c_float), ("ms_shadow", C.c_float), ("ms_shade", C.c_float), ("ms_resolve", C.c_float),
                ("ms_gather", C.c_float), ("n_gpus", C.c_uint32), ("rank", C.c_uint32),
                ("n_triangles", C.c_uint32), ("n_levels", C.c_uint32), ("n_launches", C.c_uint32), ("variant", C.c_uint32),
                ("ms_trace_primary", C.c_float), ("ms_trace_mirror", C.c_float), ("mirror_rays", C.c_uint64),
                ("thread_pencil_rays", C.c_uint64), ("ms_trace_thread", C.c_float), ("reserved", C.c_uint32)]


EXPORTS = ["rt_init", "rt_init_rank", "rt_nccl_unique_id", "rt_upload_scene", "rt_render", "rt_render_async",
           "rt_sync", "rt_download_framebuffer", "rt_download_framebuffer_u8", "rt_trace", "rt_get_stats", "rt_set_option", "rt_probe_fp32_peak",
           "rt_event_record", "rt_event_elapsed_ms", "rt_last_error", "rt_shutdown"]

_LIB = None


def lib_path():
    return os.path.join(BUILD_DIR, "librt_b200.so")


def lib():
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: the CUDA library must be built (make cuda / __graft_entry__.build()); "
                               "there is no CPU fallback")
        L = C.CDLL(path)
        L.rt_last_error.restype = C.c_char_p
        L.rt_init.argtypes = [C.c_int]
        L.rt_init_rank.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        L.rt_nccl_unique_id.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.rt_upload_scene.argtypes = [C.POINTER(RtScene)]
        L.rt_render.argtypes = [C.POINTER(RtParams)]
        L.rt_render_async.argtypes = [C.POINTER(RtParams)]
        L.rt_download_framebuffer.argtypes = [C.c_void_p, C.c_void_p]
        L.rt_download_framebuffer_u8.argtypes = [C.c_void_p]
        L.rt_trace.argtypes = [C.POINTER(RtParams), C.c_int] + [C.c_void_p] * 5
        L.rt_get_stats.argtypes = [C.POINTER(RtStats)]
        L.rt_set_option.argtypes = [C.c_int, C.c_int]
        L.rt_probe_fp32_peak.argtypes = [C.POINTER(C.c_float)]
        L.rt_event_record.argtypes = [C.c_int]
        L.rt_event_elapsed_ms.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_float)]
        L.rt_shutdown.restype = None
        _LIB = L
    return _LIB


class RtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"librt_b200 error {code}: {msg}")
        self.code = code


def _check(rc):
    if rc != 0:
        raise RtError(rc, lib().rt_last_error().decode("utf-8", "replace"))


def make_params(corners, W, H, pfx=1, pfy=None, max_lvl=10, features=RT_ALL_FEATURES, camera=(0, 0, 4), lights=None,
                want_prim_id=False):
    p = RtParams()
    c = np.ascontiguousarray(corners, np.float32).reshape(24)
    for i in range(24):
        p.corners[i] = float(c[i])
    p.width, p.height = int(W), int(H)
    p.pixelfactor_x = int(pfx)
    p.pixelfactor_y = int(pfx if pfy is None else pfy)
    p.max_lvl = int(max_lvl)
    p.features = int(features)
    cam = np.asarray(camera, np.float32)
    for i in range(3):
        p.camera[i] = float(cam[i])
    lights = np.asarray([cam] if lights is None else lights, np.float32).reshape(-1, 3)   # None: one light at the camera
    if len(lights) > RT_MAX_LIGHTS:
        raise ValueError("too many lights")
    p.n_lights = len(lights)
    for i, l in enumerate(lights):
        for k in range(3):
            p.lights[i][k] = float(l[k])
    p.want_prim_id = 1 if want_prim_id else 0
    return p


class Renderer:
    """Thin object wrapper over the C ABI (one per process; the library state is global, like the reference's)."""

    def __init__(self, n_gpus=1, device=None, rank=None, world=None, nccl_id=None):
        self.L = lib()
        if rank is None:
            _check(self.L.rt_init(int(n_gpus)))
            self.world, self.rank = int(n_gpus), 0
        else:
            buf = (C.c_char * len(nccl_id)).from_buffer_copy(nccl_id) if nccl_id else None
            _check(self.L.rt_init_rank(int(device), int(rank), int(world), buf, len(nccl_id) if nccl_id else 0))
            self.world, self.rank = int(world), int(rank)
        self._keep = None
        self.params = None

    @staticmethod
    def nccl_unique_id():
        buf = C.create_string_buffer(256)
        n = C.c_size_t(0)
        _check(lib().rt_nccl_unique_id(buf, 256, C.byref(n)))
        return bytes(buf.raw[: n.value])

    def upload_scene(self, scene):
        """scene: raytracert_b200.host.Scene (flat numpy arrays as the C++ flatten produces them)."""
        n = scene.n_triangles
        f = scene.flat()          # SoA float4 host buffers, flattened once per scene (host/flatten.h in the C++ drop-in)
        v0, v1, v2, nrm, tm = f["v0"], f["v1"], f["v2"], f["normal"], f["tri_material"]
        packed = getattr(scene, "_rt_packed", None)
        if packed is None or packed[2] is not scene.spheres or packed[3] is not scene.materials:
            mats = (RtMaterial * len(scene.materials))()
            for i, m in enumerate(scene.materials):
                for k in range(3):
                    mats[i].Kd[k], mats[i].Ka[k], mats[i].Ks[k] = float(m[k]), float(m[4 + k]), float(m[8 + k])
                mats[i].Ns, mats[i].Ni, mats[i].Tr, mats[i].flags = float(m[3]), float(m[7]), float(m[11]), int(m[12])
            sph = (RtSphere * max(1, len(scene.spheres)))()
            for i, s in enumerate(scene.spheres):
                for k in range(3):
                    sph[i].center[k] = float(s[k])
                sph[i].radius, sph[i].material = float(s[3]), int(s[4])
            packed = scene._rt_packed = (mats, sph, scene.spheres, scene.materials)
        mats, sph = packed[:2]
        sc = RtScene(n, v0.ctypes.data, v1.ctypes.data, v2.ctypes.data, nrm.ctypes.data, tm.ctypes.data,
                     len(scene.materials), C.cast(mats, C.c_void_p), len(scene.spheres),
                     C.cast(sph, C.c_void_p) if len(scene.spheres) else None)
        self._keep = (v0, v1, v2, nrm, tm, mats, sph)
        _check(self.L.rt_upload_scene(C.byref(sc)))
        self._keep = None  # the library copied everything

    def render(self, params, sync=True):
        self.params = params
        _check((self.L.rt_render if sync else self.L.rt_render_async)(C.byref(params)))

    def sync(self):
        _check(self.L.rt_sync())

    def download(self, want_prim_id=False):
        p = self.params
        rgb = np.zeros((p.height, p.width, 3), np.float32)
        prim = np.zeros(p.height * p.width * p.pixelfactor_x * p.pixelfactor_y, np.int32) if want_prim_id else None
        _check(self.L.rt_download_framebuffer(rgb.ctypes.data, prim.ctypes.data if want_prim_id else None))
        return (rgb, prim) if want_prim_id else rgb

    def download_into(self, rgb):
        _check(self.L.rt_download_framebuffer(rgb.ctypes.data, None))

    def download_u8(self):
        p = self.params
        out = np.zeros((p.height, p.width, 3), np.uint8)
        _check(self.L.rt_download_framebuffer_u8(out.ctypes.data))
        return out

    def trace(self, params, origins, dests):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dests, np.float32).reshape(-1, 3)
        n = len(o)
        rgb = np.zeros((n, 3), np.float32)
        prim = np.zeros(n, np.int32)
        hit = np.zeros((n, 3), np.float32)
        _check(self.L.rt_trace(C.byref(params), n, o.ctypes.data, d.ctypes.data, rgb.ctypes.data, prim.ctypes.data, hit.ctypes.data))
        return rgb, prim, hit

    def set_option(self, option, value):
        _check(self.L.rt_set_option(int(option), int(value)))

    def probe_fp32_peak(self):
        v = C.c_float(0)
        _check(self.L.rt_probe_fp32_peak(C.byref(v)))
        return v.value

    def stats(self):
        st = RtStats()
        _check(self.L.rt_get_stats(C.byref(st)))
        return {f[0]: getattr(st, f[0]) for f in RtStats._fields_}

    def event_record(self, slot):
        _check(self.L.rt_event_record(slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_float(0)
        _check(self.L.rt_event_elapsed_ms(a, b, C.byref(ms)))
        return ms.value

    def shutdown(self):
        self.L.rt_shutdown()
