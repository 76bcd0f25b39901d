import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np
from raytracert_b200 import binding, host, scenes
R = binding.Renderer(1)
scene = scenes.balls_standin()
cam = host.Camera(800, 800, (0.0, 2.6, 5.2), (0.0, 0.55, 0.0))
prm = binding.make_params(cam.corners, 800, 800, 4, 4, 3, 63, cam.eye, [(2.5, 4.0, 3.0)])
fb = np.zeros((800, 800, 3), np.float32)
for i in range(4):
    t0 = time.perf_counter(); R.upload_scene(scene); t1 = time.perf_counter(); R.render(prm); t2 = time.perf_counter(); R.download_into(fb); t3 = time.perf_counter()
    print(f"upload {1e3*(t1-t0):.1f} ms render {1e3*(t2-t1):.1f} ms download {1e3*(t3-t2):.1f} ms stats ms_total {R.stats()['ms_total']:.1f}")
