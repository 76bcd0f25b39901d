"""C5-style sweep on the ranks of this job: Balls stand-in at several sizes / sample counts, brute force and tile culling.
Prints one table row per configuration (rank 0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
    os.environ["NCCL_DEBUG"] = "WARN"
import numpy as np
from raytracert_b200 import binding, dist, host, scenes
R, rank, world = dist.make_renderer()
scene = scenes.balls_standin()
R.upload_scene(scene)
sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "512,1024,2048").split(",")]
pfs = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4").split(",")]
if rank == 0:
    print(f"| size | spp | GPUs | brute ms | Mrays/s | TFLOP/s alg per GPU | culled ms | Mrays/s |")
for W in sizes:
    for pf in pfs:
        cam = host.Camera(W, W, (0.0, 2.6, 5.2), (0.0, 0.55, 0.0))
        prm = binding.make_params(cam.corners, W, W, pf, pf, 3, 63, cam.eye, [(2.5, 4.0, 3.0)])
        row = []
        for cull in (0, 1):
            R.set_option(binding.RT_OPT_TILE_CULLING, cull)
            R.upload_scene(scene)
            R.render(prm)
            ms = []
            for _ in range(3):
                R.event_record(0); R.render(prm, sync=False); R.event_record(1); R.sync(); ms.append(R.event_elapsed_ms(0, 1))
            st = R.stats()
            rays = float(st["primary_rays"] + st["shadow_rays"] + st["bounce_rays"]); m = float(np.median(ms))
            if world > 1:
                import torch, torch.distributed as td
                t = torch.tensor([m], dtype=torch.float64, device="cuda"); td.all_reduce(t, op=td.ReduceOp.MAX)
                c = torch.tensor([rays], dtype=torch.float64, device="cuda"); td.all_reduce(c, op=td.ReduceOp.SUM)
                m, rays = float(t[0]), float(c[0])
            row.append((m, rays))
        if rank == 0:
            (m0, rays), (m1, _) = row
            print(f"| {W}x{W} | {pf*pf} | {world} | {m0:.2f} | {rays/m0/1e3:.1f} | {42*rays*scene.n_triangles/m0/1e9/world:.1f} | {m1:.2f} | {rays/m1/1e3:.1f} |", flush=True)
R.shutdown()
if world > 1:
    import torch.distributed as td
    td.barrier(); td.destroy_process_group()
