import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from raytracert_b200 import binding, host, scenes
R = binding.Renderer(1)
for name in ["balls", "dodge", "sphere200k"]:
    if name == "balls":
        scene = scenes.balls_standin(); cam = host.Camera(800, 800, (0.0, 2.6, 5.2), (0.0, 0.55, 0.0)); lights = [(2.5, 4.0, 3.0)]; pf = 2
    elif name == "dodge":
        scene = host.Scene.load("tests/golden/scenes/dodge.npz"); cam = host.Camera(960, 540, (.75, .55, 1.1), (.07, 0, .23)); lights = [cam.eye]; pf = 2
    else:
        scene = scenes.tessellated_sphere(400, 251); cam = host.Camera(400, 400, (0.0, 0.6, 3.4), (0, 0, 0)); lights = [(2.5, 4.0, 3.0)]; pf = 2
    R.upload_scene(scene)
    prm = binding.make_params(cam.corners, cam.W, cam.H, pf, pf, 3, 63, cam.eye, lights, want_prim_id=True)
    out = []
    for cull in (0, 1):
        R.set_option(binding.RT_OPT_TILE_CULLING, cull)
        R.upload_scene(scene)
        R.render(prm); R.render(prm)
        st = R.stats(); out.append(R.download(want_prim_id=True))
        print(f"{name:10s} cull={cull} frame {st['ms_total']:8.2f} ms trace {st['ms_trace']:8.2f} shadow {st['ms_shadow']:8.2f} exact {st['exact_evals']:.3e}", flush=True)
    print("   identical:", np.array_equal(out[0][1], out[1][1]), np.array_equal(out[0][0].view(np.uint32), out[1][0].view(np.uint32)))
