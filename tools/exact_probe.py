"""Where do the exact re-evaluations come from? Balls stand-in / dodge with feature subsets."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from raytracert_b200 import binding, host, scenes
R = binding.Renderer(1)
for name in ["balls", "dodge", "room", "sphere200k"]:
    if name == "balls":
        scene = scenes.balls_standin(); cam = host.Camera(400, 400, (0.0, 2.6, 5.2), (0.0, 0.55, 0.0)); lights = [(2.5, 4.0, 3.0)]
    elif name == "dodge":
        scene = host.Scene.load("tests/golden/scenes/dodge.npz"); cam = host.Camera(480, 270, (.75, .55, 1.1), (.07, 0, .23)); lights = [cam.eye]
    elif name == "room":
        scene = scenes.mirror_room(); cam = host.Camera(400, 400, (0.3, 1.6, 4.2), (0, 0.8, 0)); lights = [(1.5, 2.8, 2.5)]
    else:
        scene = scenes.tessellated_sphere(400, 251); cam = host.Camera(400, 400, (0.0, 0.6, 3.4), (0, 0, 0)); lights = [(2.5, 4.0, 3.0)]
    R.upload_scene(scene)
    for label, feats in [("primary only", 7), ("+shadows", 7 | 16), ("+reflection", 7 | 8), ("all", 63)]:
        prm = binding.make_params(cam.corners, cam.W, cam.H, 2, 2, 3, feats, cam.eye, lights)
        R.render(prm); R.render(prm)
        st = R.stats()
        rays = st["primary_rays"] + st["shadow_rays"] + st["bounce_rays"]
        print(f"{name:10s} {scene.n_triangles:7d} tris {label:14s} rays {rays:9d} exact {st['exact_evals']:.3e} = {st['exact_evals'] / rays:8.1f}/ray ({st['exact_evals'] / st['tri_tests']:.2e}/test)  "
              f"ms trace {st['ms_trace']:.2f} shadow {st['ms_shadow']:.2f}  {st['tri_tests'] / (st['ms_trace'] + st['ms_shadow']) / 1e9:.0f} Gtests/s", flush=True)
