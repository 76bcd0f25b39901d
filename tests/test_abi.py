"""The drop-in boundary: librt_b200.so loads without a GPU, exports every entry point include/rt_b200.h
declares, the ctypes structs have the header's layout, and without a device every call fails loudly
(there is no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "rt_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol(built):
    from raytracert_b200 import binding
    names = declared_functions()
    assert {"rt_upload_scene", "rt_render", "rt_download_framebuffer", "rt_trace", "rt_init"} <= set(names)
    L = C.CDLL(binding.lib_path())
    for n in names:
        assert hasattr(L, n), f"{n} is declared in rt_b200.h but not exported"
    assert sorted(binding.EXPORTS) == names, "binding.EXPORTS out of date with the header"


def test_no_oracle_or_torch_in_product_library(built):
    """The product library must not link the oracle (or anything but libc/libstdc++/libdl-class system libs)."""
    from raytracert_b200 import binding
    out = subprocess.run(["ldd", binding.lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out and "torch" not in out and "libcudart" not in out  # cudart is linked statically
    syms = subprocess.run(["nm", "-D", "--defined-only", binding.lib_path()], capture_output=True, text=True).stdout
    assert "orc_" not in syms and "ref_render" not in syms


def test_struct_layouts_match_header(built, tmp_path):
    from raytracert_b200 import binding
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "rt_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(rt_material),sizeof(rt_sphere),sizeof(rt_scene),sizeof(rt_params),sizeof(rt_stats),"
                   "offsetof(rt_params,lights),offsetof(rt_params,want_prim_id),offsetof(rt_stats,ms_total));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(binding.RtMaterial), C.sizeof(binding.RtSphere), C.sizeof(binding.RtScene), C.sizeof(binding.RtParams),
            C.sizeof(binding.RtStats), binding.RtParams.lights.offset, binding.RtParams.want_prim_id.offset, binding.RtStats.ms_total.offset]
    assert got == want


def test_option_and_flag_constants_match_header(built, tmp_path):
    """The ctypes glue repeats the header's option numbers and feature / material flag bits: keep them in step."""
    from raytracert_b200 import binding
    src = tmp_path / "opt.c"
    src.write_text('#include <stdio.h>\n#include "rt_b200.h"\nint main(void){printf("%d %d %d %d %d %d %d %d %d\\n",'
                   "RT_OPT_TILE_CULLING,RT_OPT_PENCIL,RT_OPT_PENCIL_ANY,RT_OPT_GRAPH,RT_OPT_PENCIL_REFLECT,RT_OPT_SMALL_TRACE,RT_OPT_PENCIL_THREAD,RT_MAX_LIGHTS,"
                   "RT_AMBIENT|RT_DIFFUSE|RT_SPECULAR|RT_REFLECTION|RT_SHADOWS|RT_REFRACTION);return 0;}\n")
    exe = tmp_path / "opt"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [binding.RT_OPT_TILE_CULLING, binding.RT_OPT_PENCIL, binding.RT_OPT_PENCIL_ANY, binding.RT_OPT_GRAPH, binding.RT_OPT_PENCIL_REFLECT, binding.RT_OPT_SMALL_TRACE, binding.RT_OPT_PENCIL_THREAD, binding.RT_MAX_LIGHTS, binding.RT_ALL_FEATURES]


def test_header_is_plain_c(built, tmp_path):
    src = tmp_path / "c.c"
    src.write_text('#include "rt_b200.h"\nint main(void){return RT_OK;}\n')
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "c.o")], check=True)


def test_fails_loudly_without_a_device(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from raytracert_b200 import binding
    L = binding.lib()
    assert L.rt_init(1) == -1                       # RT_ERR_NO_DEVICE
    assert b"no CPU fallback" in L.rt_last_error()
    p = binding.make_params([0] * 24, 4, 4)
    assert L.rt_render(C.byref(p)) == -3            # RT_ERR_STATE: nothing was initialised, nothing is rendered
    with pytest.raises(binding.RtError):
        binding.Renderer(1)
