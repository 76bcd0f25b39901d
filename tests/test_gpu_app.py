"""The C++ drop-in (raytracert_b200/host/main.cpp: the reference's skeleton without GLUT) end to end on the GPU:
OBJ/MTL load -> rt_upload_scene -> 'r' key -> rt_render -> Image::writeImage, compared with the PPM the
unmodified reference would have written for the same scene, camera and keys (fixture quirks_72_pf2)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_case

pytestmark = pytest.mark.gpu
APP = os.path.join(ROOT, "raytracert_b200", "_build", "rt_main")


def read_ppm(path):
    raw = open(path, "rb").read()
    assert raw[:3] == b"P6\n"
    parts = raw.split(b"\n", 3)
    w, h = (int(x) for x in parts[1].split())
    assert parts[2] == b"255"
    return np.frombuffer(parts[3], np.uint8).reshape(h, w, 3)


def run_app(tmp_path, *args):
    if not os.path.exists(APP):
        subprocess.run(["make", "-C", ROOT, "app"], check=True, stdout=subprocess.DEVNULL)
    out = str(tmp_path / "result.ppm")
    r = subprocess.run([APP, os.path.join(GOLDEN, "obj", "quirks.obj"), "--out", out, *args], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:]
    return read_ppm(out), r.stdout


def test_app_writes_the_reference_image(built, tmp_path):
    c = load_case("quirks_72_pf2")
    img, log = run_app(tmp_path, "--size", "72x72", "--pf", "2", "--lvl", "6", "--eye", "3.4,3.0,4.6", "--center", "0.4,0.2,0.2",
                       "--light", "3,5,4")
    assert img.shape == c["u8"].shape
    d = np.abs(img.astype(int) - c["u8"].astype(int))
    assert d.max() <= 1 and np.mean(np.any(d > 0, axis=2)) <= 0.01
    assert "Raytracing" in log and "Mrays/s" in log


def test_app_key_replay_toggles(built, tmp_path, port):
    """--keys replays the reference's key presses: '5' toggles shadows off, '4' reflection off, then 'r' renders."""
    from conftest import load_scene
    from raytracert_b200 import host
    img, _ = run_app(tmp_path, "--size", "64x48", "--pf", "1", "--lvl", "6", "--eye", "3.4,3.0,4.6", "--center", "0.4,0.2,0.2",
                     "--light", "3,5,4", "--keys", "54r")
    cam = host.Camera(64, 48, (3.4, 3.0, 4.6), (0.4, 0.2, 0.2))
    port.set_scene(load_scene("quirks"))
    port.configure(cam.eye, [(3, 5, 4)], 63 & ~16 & ~8, 6)
    rgb, _, _ = port.render(cam.corners, 64, 48, 1, 1)
    d = np.abs(img.astype(int) - port.quantise(rgb).astype(int))
    assert d.max() <= 1


def test_app_light_and_sample_keys(built, tmp_path, port):
    """'L' adds a light at the camera (main.cpp:334-336), '+' raises pixelfactorX/Y (raytracing.cpp:474-477); no --light:
    init() puts the first light at the start-up camera position (raytracing.cpp:72)."""
    from conftest import load_scene
    from raytracert_b200 import host
    img, _ = run_app(tmp_path, "--size", "48x40", "--pf", "1", "--lvl", "3", "--eye", "3.4,3.0,4.6", "--center", "0.4,0.2,0.2", "--keys", "L+r")
    cam = host.Camera(48, 40, (3.4, 3.0, 4.6), (0.4, 0.2, 0.2))
    port.set_scene(load_scene("quirks"))
    port.configure(cam.eye, [cam.eye, cam.eye], 63, 3)
    rgb, _, _ = port.render(cam.corners, 48, 40, 2, 2)
    d = np.abs(img.astype(int) - port.quantise(rgb).astype(int))
    assert d.max() <= 1


def _device_count():
    r = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True)
    return sum(1 for l in r.stdout.splitlines() if l.startswith("GPU ")) if r.returncode == 0 else 0


def test_single_process_multi_gpu_app(built, tmp_path):
    """rt_init(n) -- one process drives n GPUs (ncclCommInitAll), what `rt_main --gpus N` and INTEGRATION.md use: rows
    interleaved over the devices, one all-gather, de-interleave.  The PPM must be byte-identical to the --gpus 1 one.
    Skips on a box with one GPU (the driver's SCALE run and tools/gpu scripts exercise it at 2/4/8)."""
    n = _device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    args = ["--size", "96x70", "--pf", "2", "--lvl", "6", "--eye", "3.4,3.0,4.6", "--center", "0.4,0.2,0.2", "--light", "3,5,4"]
    one, _ = run_app(tmp_path, *args, "--gpus", "1")
    for g in sorted({2, min(n, 4), n}):
        many, _ = run_app(tmp_path, *args, "--gpus", str(g))
        assert np.array_equal(one, many), f"--gpus {g}: {int(np.count_nonzero(one != many))} bytes differ"


def test_single_process_multi_gpu_ids(built, tmp_path):
    """The same through the C ABI from Python in a fresh process: Renderer(n) against Renderer(1), float framebuffer bits
    and per-sample ids (rt_download_framebuffer gathers the ids of every device in single-process mode)."""
    n = _device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from raytracert_b200 import binding, host, scenes\n"
        "s = scenes.balls_standin(grid=48, slices=24, stacks=12)\n"
        "cam = host.Camera(120, 90, (0.0, 2.6, 5.2), (0.0, 0.55, 0.0))\n"
        "prm = binding.make_params(cam.corners, 120, 90, 2, 2, 3, 63, cam.eye, [(2.5, 4.0, 3.0)], want_prim_id=True)\n"
        "out = {}\n"
        f"for g in (1, 2, {n}):\n"
        "    R = binding.Renderer(g); R.upload_scene(s); R.render(prm); out[g] = R.download(want_prim_id=True); st = R.stats(); R.shutdown()\n"
        "    assert st['n_gpus'] == g\n"
        "    assert np.array_equal(out[1][1], out[g][1]), g\n"
        "    assert np.array_equal(out[1][0].view(np.uint32), out[g][0].view(np.uint32)), g\n"
        "print('OK')\n")
    r = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


def test_trace_and_stats_before_any_frame(built):
    """A fresh process that never renders a frame: rt_trace, rt_get_stats (which queries frame-phase events nobody recorded),
    rt_trace again with and without the graph replay -- the stale CUDA error code of the event query must not surface."""
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from raytracert_b200 import binding, host\n"
        f"z = np.load({os.path.join(GOLDEN, 'trace_shadow_test.npz')!r})\n"
        f"s = host.Scene.load({os.path.join(GOLDEN, 'scenes', 'shadow_test.npz')!r})\n"
        "R = binding.Renderer(1); R.upload_scene(s)\n"
        "prm = binding.make_params([0] * 24, 1, 1, 1, 1, 10, 63, z['eye'], [z['eye']])\n"
        "for g in (0, -1, 0, 1):\n"
        "    R.set_option(binding.RT_OPT_GRAPH, g)\n"
        "    for n in (1, 7, 1):\n"
        "        rgb, prim, hit = R.trace(prm, z['origins'][:n], z['dests'][:n])\n"
        "        assert np.array_equal(prim, z['prim'][:n]), (g, n)\n"
        "        st = R.stats()\n"
        "        assert bool(st['variant'] & 16) == (g != 0), (g, st['variant'])\n"
        "R.shutdown(); print('OK')\n")
    r = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
