"""Sweep the compiled scan-kernel shapes (RT_B200_TUNE) on cuda:0: Balls stand-in, device time per kernel kind.
usage: python tools/tune.py [W] [pf] [configs "rp,j,minb;..."]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from raytracert_b200 import binding, host, scenes

W = int(sys.argv[1]) if len(sys.argv) > 1 else 800
pf = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfgs = sys.argv[3].split(";") if len(sys.argv) > 3 else ["2,8,2", "2,8,3", "1,8,4", "1,16,4", "1,8,3", "1,16,3", "1,8,5", "1,16,5", "2,4,2", "2,4,3"]
scene = scenes.balls_standin()
cam = host.Camera(W, W, (0.0, 2.6, 5.2), (0.0, 0.55, 0.0))
prm = binding.make_params(cam.corners, W, W, pf, pf, 3, 63, cam.eye, [(2.5, 4.0, 3.0)], want_prim_id=True)
base = None
for c in cfgs:
    os.environ["RT_B200_TUNE"] = c
    R = binding.Renderer(1)
    R.upload_scene(scene)
    R.render(prm)
    ms = []
    for _ in range(3):
        R.event_record(0); R.render(prm, sync=False); R.event_record(1); R.sync()
        ms.append(R.event_elapsed_ms(0, 1))
    st = R.stats()
    rgb, prim = R.download(want_prim_id=True)
    if base is None:
        base = (rgb, prim)
    same = np.array_equal(prim, base[1]) and np.array_equal(rgb.view(np.uint32), base[0].view(np.uint32))
    tests = st["tri_tests"]
    print(f"{c:8s} frame {min(ms):8.2f} ms  trace {st['ms_trace']:8.2f}  shadow {st['ms_shadow']:8.2f}  shade {st['ms_shade']:.2f}  "
          f"{tests / min(ms) / 1e9:7.1f} Gtests/s  exact {st['exact_evals']:.3e}  identical_to_first={same}", flush=True)
    R.shutdown()
