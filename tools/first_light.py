"""Scratch driver: render a few scenes on cuda:0 and compare with the port oracle (bitwise ids, RGB tolerance)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from raytracert_b200 import host, scenes, binding
from oracle import pyoracle

def compare(name, R, P, scene, cam, pf, maxlvl, lights, feats=63):
    W, H = cam.W, cam.H
    lights = [cam.eye] if lights is None else lights
    P.set_scene(scene); P.configure(cam.eye, lights, feats, maxlvl); P.reset_counts()
    t = time.time(); rgb_o, srgb_o, prim_o = P.render(cam.corners, W, H, pf, pf, want_samples=True); t_cpu = time.time() - t
    counts = P.ray_counts()
    R.upload_scene(scene)
    prm = binding.make_params(cam.corners, W, H, pf, pf, maxlvl, feats, cam.eye, lights, want_prim_id=True)
    R.render(prm)
    t = time.time(); R.render(prm); t_gpu = time.time() - t
    rgb_g, prim_g = R.download(want_prim_id=True)
    st = R.stats()
    idmis = int(np.sum(prim_o != prim_g))
    d = np.abs(rgb_o - rgb_g)
    u8o, u8g = P.quantise(rgb_o).astype(int), R.download_u8().astype(int)
    bad = np.mean(np.any(np.abs(u8o - u8g) > 1, axis=2))
    print(f"{name:34s} ids_mismatch={idmis}/{prim_o.size} maxabs={d.max():.2e} px>1/255={bad:.5f} rays cpu={counts} gpu=({st['primary_rays']},{st['shadow_rays']},{st['bounce_rays']}) exact={st['exact_evals']} ({st['exact_evals']/max(1,st['tri_tests']):.2e}/test) gpu={t_gpu*1e3:.1f}ms cpu={t_cpu:.2f}s", flush=True)

R = binding.Renderer(1)
P = pyoracle.PortOracle()
cube = scenes.unit_cube()
room = scenes.mirror_room()
balls = scenes.balls_standin(grid=48, slices=24, stacks=12)
compare("cube default 128", R, P, cube, host.Camera(128, 128), 1, 10, None)
compare("cube oblique 160 pf2", R, P, cube, host.Camera(160, 160, (2.6, 2.4, 3.0), (.5, .5, .5)), 2, 10, None)
compare("room 128 pf2 lvl4", R, P, room, host.Camera(128, 128, (0.3, 1.6, 4.2), (0, 0.8, 0)), 2, 4, [(1.5, 2.8, 2.5)])
compare("room 96 2 lights lvl10", R, P, room, host.Camera(96, 96, (0.3, 1.6, 4.2), (0, 0.8, 0)), 1, 10, [(1.5, 2.8, 2.5), (-1.0, 2.0, 1.0)])
compare("balls-small 128 pf2 lvl3", R, P, balls, host.Camera(128, 128, (0.0, 2.6, 5.2), (0.0, 0.55, 0.0)), 2, 3, [(2.5, 4.0, 3.0)])
for f in ["dodge", "shadow_test"]:
    p = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", f + ".npz")
    if os.path.exists(p):
        sc = host.Scene.load(p)
        cam = host.Camera(96, 54, (.75, .55, 1.1), (.07, 0, .23)) if f == "dodge" else host.Camera(96, 96, (1, 5, 7), (1, 1.2, .7))
        compare(f, R, P, sc, cam, 1, 10, None)
# throughput probe
big = scenes.balls_standin()
R.upload_scene(big)
cam = host.Camera(800, 800, (0.0, 2.6, 5.2), (0.0, 0.55, 0.0))
prm = binding.make_params(cam.corners, 800, 800, 2, 2, 3, 63, cam.eye, [(2.5, 4.0, 3.0)])
R.render(prm)
R.event_record(0); R.render(prm, sync=False); R.event_record(1); R.sync()
ms = R.event_elapsed_ms(0, 1); st = R.stats()
print(f"balls 800x800 pf2 lvl3: {ms:.1f} ms, rays={st['primary_rays']+st['shadow_rays']+st['bounce_rays']}, tests={st['tri_tests']:.3e}, {st['tri_tests']/ms/1e9:.1f} Gtests/s -> alg {42*st['tri_tests']/ms/1e9/1e3:.1f} TFLOP/s; exact={st['exact_evals']}")
