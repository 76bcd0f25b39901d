"""Run one BASELINE configuration end to end (single GPU, or under torchrun with rows interleaved over the ranks):
one warm-up frame, `--frames` timed frames per mode (brute force, then opt-in tile culling), an oracle check on a
pixel lattice (rank 0), one JSON line.  Used for the configurations that are too heavy for bench.py's default loop
(C4: 1 M triangles at 3840x2160x16).

  python -m torch.distributed.run --nproc-per-node 8 ... tools/run_config.py --workload sphere1m --frames 1
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
    os.environ["NCCL_DEBUG"] = "WARN"
import numpy as np
import bench
from raytracert_b200 import binding, dist, host

bench.guard_stdout()   # NCCL's banner and friends go to stderr; stdout carries the one JSON line only

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="sphere1m")
ap.add_argument("--frames", type=int, default=1)
ap.add_argument("--lattice", type=int, default=0, help="oracle check on every k-th pixel of every k-th row (0 = ~150 pixels)")
ap.add_argument("--modes", default="brute_force,tile_culling", help="comma list of brute_force, tile_culling")
ap.add_argument("--cpu-pixels", type=float, default=0, help="> 0: also time the CPU reference on a lattice of about this many pixels (all cores + 1 thread)")
args = ap.parse_args()
R, rank, world = dist.make_renderer()
scene, W, H, pf, lvl, eye, center, lights, desc = bench.workload(args.workload)
cam = host.Camera(W, H, eye, center)
lights = [cam.eye] if lights is None else lights
R.upload_scene(scene)
small = binding.make_params(host.Camera(64, 36, eye, center).corners, 64, 36, 1, 1, lvl, 63, cam.eye, lights)
prm = binding.make_params(cam.corners, W, H, pf, pf, lvl, 63, cam.eye, lights, want_prim_id=(world == 1))
out = {"workload": desc, "n_gpus": world}
frames = {}
modes = args.modes.split(",")
for mode, cull in (("brute_force", 0), ("tile_culling", 1)):
    if mode not in modes:
        continue
    R.set_option(binding.RT_OPT_TILE_CULLING, cull)
    R.upload_scene(scene)
    # warm-up (allocations, records, NCCL connections): the full frame when it is cheap, a tiny one otherwise
    R.render(prm if W * H * pf * pf * float(scene.n_triangles) < 2e11 else small)
    ms = []
    for _ in range(args.frames):
        R.event_record(0); R.render(prm, sync=False); R.event_record(1); R.sync()
        ms.append(R.event_elapsed_ms(0, 1))
    st = R.stats()
    rays = st["primary_rays"] + st["shadow_rays"] + st["bounce_rays"]
    if world > 1:
        import torch, torch.distributed as td
        t = torch.tensor([float(np.mean(ms))], dtype=torch.float64, device="cuda"); td.all_reduce(t, op=td.ReduceOp.MAX)
        c = torch.tensor([float(rays)], dtype=torch.float64, device="cuda"); td.all_reduce(c, op=td.ReduceOp.SUM)
        m, rays = float(t[0]), float(c[0])
    else:
        m = float(np.mean(ms))
    frames[mode] = R.download(want_prim_id=(world == 1))
    out[mode] = {"ms_per_frame": m, "Mrays_per_s": rays / m / 1e3, "rays": rays, "tests_per_s": rays * scene.n_triangles / (m * 1e-3), "variant": st["variant"]}
    if not cull:
        # executed FP32 at the pipe: hot-loop flops per test (12 pencil / 27 generic) x this rank's tests of each launch kind
        v = st["variant"]
        cnt = np.array([st["primary_rays"], st["bounce_rays"], st["shadow_rays"], st["mirror_rays"], st["thread_pencil_rays"]], np.float64)
        kinds = np.array([st["ms_trace_primary"], st["ms_trace"] - st["ms_trace_primary"], st["ms_shadow"]], np.float64)
        if world > 1:
            c = torch.tensor(cnt, dtype=torch.float64, device="cuda"); td.all_reduce(c, op=td.ReduceOp.SUM); cnt = c.cpu().numpy()
            k_ = torch.tensor(kinds, dtype=torch.float64, device="cuda"); td.all_reduce(k_, op=td.ReduceOp.MAX); kinds = k_.cpu().numpy()
        peak = 148 * 128 * 2 * 1.965e9 / 1e12
        alg = 42 * rays * scene.n_triangles / (m * 1e-3) / 1e12 / world
        # (bounce rays: generic 27, served by a mirror pencil 12, by thread pencils 23.6)
        ex = ((12 if v & 2 else 27) * cnt[0] + 27 * (cnt[1] - cnt[3] - cnt[4]) + 12 * cnt[3] + 23.6 * cnt[4] + (12 if v & 4 else 27) * cnt[2]) \
            * scene.n_triangles / (m * 1e-3) / 1e12 / world
        out[mode].update({"fp32_algorithmic_tflops_per_gpu": alg, "algorithmic_ratio": alg / peak, "fp32_executed_tflops_per_gpu": ex,
                          "executed_frac_of_fp32_peak": ex / peak, "ms_primary_bounce_shadow": [float(x) for x in kinds],
                          "rays_primary_bounce_shadow_mirror_thread": [float(x) for x in cnt]})
if rank == 0:
    a = frames[modes[0]]
    rgb = a[0] if world == 1 else a
    if len(frames) == 2:
        b = frames["tile_culling"]
        rgb_c = b[0] if world == 1 else b
        out["culling_bit_identical"] = bool(np.array_equal(rgb.view(np.uint32), rgb_c.view(np.uint32)))
    from oracle import pyoracle
    k = args.lattice or max(1, int(round((W * H / 150.0) ** 0.5)))
    P = pyoracle.PortOracle(); P.set_scene(scene); P.configure(cam.eye, lights, 63, lvl)
    t0 = time.time()
    rgb_o, _, prim_o = P.render(cam.corners, W, H, pf, pf, y0=k // 2, ystep=k, x0=k // 2, xstep=k, want_samples=True)
    ys, xs = np.arange(k // 2, H, k), np.arange(k // 2, W, k)
    d = np.abs(rgb[np.ix_(ys, xs)] - rgb_o[np.ix_(ys, xs)])
    out["oracle_lattice"] = {"pixels": int(len(ys) * len(xs)), "max_abs_rgb_diff": float(d.max()), "cpu_seconds": time.time() - t0,
                             "hit_pixels": int(np.count_nonzero(rgb_o[np.ix_(ys, xs)].sum(axis=2) > 0))}
    if world == 1:
        po = prim_o.reshape(H, W, pf * pf)[np.ix_(ys, xs)]; pg = a[1].reshape(H, W, pf * pf)[np.ix_(ys, xs)]
        out["oracle_lattice"]["id_mismatches"] = int(np.count_nonzero(po != pg))
    if args.cpu_pixels > 0:
        out["cpu_baseline"] = bench.cpu_baseline_block(scene, cam, pf, lvl, lights, os.cpu_count() or 1, args.cpu_pixels)
    bench.emit(out)
R.shutdown()
if world > 1:
    import torch.distributed as td
    td.barrier(); td.destroy_process_group()
