"""Two frames of the Balls stand-in at 800x800, 1 ray/pixel (for ncu: -k regex:k_trace -s 4 -c 1 = the 2nd frame's primary scan)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raytracert_b200 import binding, host, scenes
scene = scenes.balls_standin()
cam = host.Camera(800, 800, (0.0, 2.6, 5.2), (0.0, 0.55, 0.0))
pf = int(sys.argv[1]) if len(sys.argv) > 1 else 1
prm = binding.make_params(cam.corners, 800, 800, pf, pf, 3, 63, cam.eye, [(2.5, 4.0, 3.0)])
R = binding.Renderer(1); R.upload_scene(scene)
for _ in range(2):
    R.render(prm)
st = R.stats()
print(st["ms_trace"], st["ms_shadow"], st["exact_evals"])
R.shutdown()
