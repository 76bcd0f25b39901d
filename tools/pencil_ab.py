"""A/B of RT_OPT_PENCIL on one workload (single GPU): renders the frame with the generic filter and with the pencil
filter, checks that ids and framebuffer bits are identical, prints per-kernel times and exact re-evaluation counts.

  python tools/pencil_ab.py --workload balls --frames 3
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from raytracert_b200 import binding, host

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="balls")
ap.add_argument("--frames", type=int, default=3)
args = ap.parse_args()
scene, W, H, pf, lvl, eye, center, lights, desc = bench.workload(args.workload)
cam = host.Camera(W, H, eye, center)
lights = [cam.eye] if lights is None else lights
R = binding.Renderer(1)
R.upload_scene(scene)
prm = binding.make_params(cam.corners, W, H, pf, pf, lvl, 63, cam.eye, lights, want_prim_id=True)
out = {"workload": desc}
frames = {}
for name, opt in (("generic", 0), ("pencil", 1)):
    R.set_option(binding.RT_OPT_PENCIL, opt)
    R.render(prm)
    ms = []
    for _ in range(args.frames):
        R.event_record(0); R.render(prm, sync=False); R.event_record(1); R.sync()
        ms.append(R.event_elapsed_ms(0, 1))
    st = R.stats()
    frames[name] = R.download(want_prim_id=True)
    out[name] = {"ms_per_frame": float(np.median(ms)), "ms_trace": st["ms_trace"], "ms_shadow": st["ms_shadow"], "exact_evals": st["exact_evals"],
                 "variant": st["variant"], "n_launches": st["n_launches"], "rays": st["primary_rays"] + st["shadow_rays"] + st["bounce_rays"]}
a, b = frames["generic"], frames["pencil"]
out["ids_identical"] = bool(np.array_equal(a[1], b[1]))
out["rgb_bits_identical"] = bool(np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)))
out["speedup"] = out["generic"]["ms_per_frame"] / out["pencil"]["ms_per_frame"]
print(json.dumps(out))
R.shutdown()
