"""C5 sweep (BASELINE configs[4]): the Balls stand-in at several frame sizes and sample counts on the ranks of this job
(1 GPU, or under torchrun with rows interleaved over the ranks), brute force and opt-in tile culling.  One JSON object
per configuration (rank 0, stdout), e.g.

  python tools/sweep.py --sizes 512,1024,2048 --pf 1,2,4 > profiles/r2_C5_n1.jsonl
  python -m torch.distributed.run --nproc-per-node 8 ... tools/sweep.py --sizes 8192 --pf 8 --frames 1

The pixelfactor is the reference's own supersampling knob (raytracing.cpp:23-25, '+'/'-' keys :475-477; loop bounds
main.cpp:360-362).  Every number is device time (CUDA events on the library's stream), max over ranks."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
    os.environ["NCCL_DEBUG"] = "WARN"
import numpy as np
import bench
from raytracert_b200 import binding, dist, host, scenes

bench.guard_stdout()   # NCCL's banner and friends go to stderr; stdout carries the JSON rows only

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="512,1024,2048")
ap.add_argument("--pf", default="1,2,4")
ap.add_argument("--frames", type=int, default=2)
ap.add_argument("--max-samples", type=float, default=1e12, help="skip configurations with more samples than this per GPU")
ap.add_argument("--min-samples", type=float, default=0.0)
ap.add_argument("--no-cull", action="store_true")
args = ap.parse_args()
R, rank, world = dist.make_renderer()
scene = scenes.balls_standin()
PEAK = 148 * 128 * 2 * 1.965e9 / 1e12
for W in [int(x) for x in args.sizes.split(",")]:
    for pf in [int(x) for x in args.pf.split(",")]:
        samples = float(W) * W * pf * pf
        if samples / world > args.max_samples or samples / world < args.min_samples:
            continue
        cam = host.Camera(W, W, (0.0, 2.6, 5.2), (0.0, 0.55, 0.0))
        prm = binding.make_params(cam.corners, W, W, pf, pf, 3, 63, cam.eye, [(2.5, 4.0, 3.0)])
        warm = binding.make_params(host.Camera(64, 64, (0.0, 2.6, 5.2), (0.0, 0.55, 0.0)).corners, 64, 64, 1, 1, 3, 63, cam.eye, [(2.5, 4.0, 3.0)])
        row = {"config": "C5", "size": W, "spp": pf * pf, "n_gpus": world, "triangles": scene.n_triangles}
        for mode, cull in (("brute_force", 0), ("tile_culling", 1)):
            if cull and args.no_cull:
                continue
            R.set_option(binding.RT_OPT_TILE_CULLING, cull)
            R.upload_scene(scene)
            R.render(prm if samples / world < 5e7 else warm)     # warm-up: the frame itself when it is cheap
            ms = []
            for _ in range(args.frames):
                R.event_record(0); R.render(prm, sync=False); R.event_record(1); R.sync(); ms.append(R.event_elapsed_ms(0, 1))
            st = R.stats()
            rays = float(st["primary_rays"] + st["shadow_rays"] + st["bounce_rays"]); m = float(np.min(ms))
            kinds = [st["ms_trace_primary"], st["ms_trace"] - st["ms_trace_primary"], st["ms_shadow"]]
            cnt = [float(st["primary_rays"]), float(st["bounce_rays"]), float(st["shadow_rays"]), float(st["mirror_rays"]), float(st["thread_pencil_rays"])]
            if world > 1:
                import torch, torch.distributed as td
                t = torch.tensor([m] + kinds, dtype=torch.float64, device="cuda"); td.all_reduce(t, op=td.ReduceOp.MAX)
                c = torch.tensor([rays] + cnt, dtype=torch.float64, device="cuda"); td.all_reduce(c, op=td.ReduceOp.SUM)
                m, kinds, rays, cnt = float(t[0]), [float(x) for x in t[1:]], float(c[0]), [float(x) for x in c[1:]]
            alg = 42 * rays * scene.n_triangles / (m * 1e-3) / 1e12 / world
            row[mode] = {"ms_per_frame": m, "Mrays_per_s": rays / m / 1e3, "rays": rays, "variant": st["variant"]}
            if not cull:
                # executed FP32 at the pipe: hot-loop flops per test (12 pencil / 27 generic) x tests of each launch kind
                pencil_p, pencil_s = bool(st["variant"] & 2), bool(st["variant"] & 4)
                # (bounce rays: generic 27, served by a mirror pencil 12, by thread pencils 23.6)
                ex = ((12 if pencil_p else 27) * cnt[0] + 27 * (cnt[1] - cnt[3] - cnt[4]) + 12 * cnt[3] + 23.6 * cnt[4] + (12 if pencil_s else 27) * cnt[2]) \
                    * scene.n_triangles / (m * 1e-3) / 1e12 / world
                row[mode].update({"fp32_algorithmic_tflops_per_gpu": alg, "algorithmic_ratio": alg / PEAK, "fp32_executed_tflops_per_gpu": ex,
                                  "executed_frac_of_fp32_peak": ex / PEAK, "ms_primary_bounce_shadow": kinds})
        if rank == 0:
            bench.emit(row)
R.set_option(binding.RT_OPT_TILE_CULLING, 0)
R.shutdown()
if world > 1:
    import torch.distributed as td
    td.barrier(); td.destroy_process_group()
