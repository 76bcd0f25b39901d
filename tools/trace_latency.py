"""Latency of the drop-in performRayTracing(origin, dest) call = rt_trace with n = 1 (and small batches), with and
without the CUDA-graph replay (RT_OPT_GRAPH), on shadow_test (the scene of BASELINE.md's CPU probe: 0.08 ms per call on
the CPU) and on the headline scene.  Prints one JSON line.   python tools/trace_latency.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from raytracert_b200 import binding, host, scenes

R = binding.Renderer(1)
out = {}
z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "trace_shadow_test.npz"))
for name, scene, eye in (("shadow_test", host.Scene.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", "shadow_test.npz")), z["eye"]),
                         ("balls", scenes.balls_standin(), np.array([0.0, 2.6, 5.2], np.float32))):
    R.upload_scene(scene)
    o, d = z["origins"], z["dests"]
    for lvl in (10, 3):
        prm = binding.make_params([0] * 24, 1, 1, 1, 1, lvl, 63, eye, [eye])
        for graph in (0, -1):
            R.set_option(binding.RT_OPT_GRAPH, graph)
            for n in (1, 32):
                for i in range(20):
                    R.trace(prm, o[i:i + n], d[i:i + n])
                t = time.perf_counter()
                reps = 300
                for i in range(reps):
                    R.trace(prm, o[i % 500:i % 500 + n], d[i % 500:i % 500 + n])
                dt = (time.perf_counter() - t) / reps
                st = R.stats()
                out[f"{name} ({scene.n_triangles} tri) lvl{lvl} n={n} graph={'auto' if graph < 0 else 'off'}"] = {
                    "ms_per_call": dt * 1e3, "launches": st["n_launches"], "graph_replay": bool(st["variant"] & 16)}
print(json.dumps(out, indent=1))
R.shutdown()
