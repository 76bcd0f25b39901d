import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from raytracert_b200 import binding, host, scenes
scene = scenes.balls_standin()
cam = host.Camera(800, 800, (0.0, 2.6, 5.2), (0.0, 0.55, 0.0))
prm = binding.make_params(cam.corners, 800, 800, 2, 2, 3, 63, cam.eye, [(2.5, 4.0, 3.0)], want_prim_id=True)
base = None
for cfg in ["2,8,2", "1,16,4", "1,8,4"]:
  for cm in ["1e-3", "3e-4", "1e-4", "3e-5", "1e-5"]:
    os.environ["RT_B200_TUNE"] = cfg; os.environ["RT_B200_COSMIN"] = cm
    R = binding.Renderer(1); R.upload_scene(scene); R.render(prm)
    ms = []
    for _ in range(3):
        R.event_record(0); R.render(prm, sync=False); R.event_record(1); R.sync(); ms.append(R.event_elapsed_ms(0, 1))
    st = R.stats(); rgb, prim = R.download(want_prim_id=True)
    if base is None: base = (rgb, prim)
    same = np.array_equal(prim, base[1]) and np.array_equal(rgb.view(np.uint32), base[0].view(np.uint32))
    print(f"{cfg:7s} cosmin {cm:5s} frame {min(ms):8.2f} ms trace {st['ms_trace']:8.2f} shadow {st['ms_shadow']:8.2f} exact {st['exact_evals']:.3e} same={same}", flush=True)
    R.shutdown()
