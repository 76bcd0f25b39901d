#!/usr/bin/env python
"""bench.py -- headline benchmark of the render hot path (BASELINE.json): Mrays/s and ms/frame on the Balls
stand-in scene, 800x800, 4x4 rays/pixel, shadows + reflection depth 3, at N GPUs of one box.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]
  torchrun --nproc-per-node N ... bench.py --gpus N ...     (one rank per GPU; rows interleaved over ranks)

A "step" is one frame: rt_render through the C ABI of librt_b200.so (ray generation, nearest-hit scans,
shadow scans, shading, reflection bounces, resolve, and for N > 1 the NCCL all-gather of the row slabs).
`value` counts every ray the reference would cast (intersectMesh calls: primary + shadow + continuation).
Prints ONE JSON line (rank 0).  `--impl reference` times the reference's own CPU code (oracle/_ref when it
was built from /root/reference, else the plain-C port) on the host cores, on a bounded sample of rows.

The oracle is used here only as the CPU baseline / reference arm -- never on the product path.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: keep NCCL's version banner (printed at NCCL_DEBUG=VERSION and above) off it
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
    os.environ["NCCL_DEBUG"] = "WARN"

# ... and whatever a native library still prints to file descriptor 1 (NCCL's banner at NCCL_DEBUG=INFO set in a config
# file, for one) must not reach it either: fd 1 is pointed at stderr for the whole run, the JSON line goes to the saved fd.
_JSON_FD = None


def guard_stdout():
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    if _JSON_FD is None:
        print(json.dumps(line), flush=True)
    else:
        os.write(_JSON_FD, (json.dumps(line) + "\n").encode())


METRIC = "Mrays/s (Balls stand-in 800x800, 16 spp, shadows + reflection depth 3)"
FLOPS_PER_TEST = 42  # minimal ray-dependent restatement of raytracing.cpp:111-151, FMA = 2


def workload(name):
    """(scene, W, H, pf, max_lvl, look-at eye, center, lights, description)"""
    from raytracert_b200 import host, scenes
    if name == "balls":      # BASELINE configs[1] / the metric's configuration
        s = scenes.balls_standin()
        return s, 800, 800, 4, 3, (0.0, 2.6, 5.2), (0.0, 0.55, 0.0), [(2.5, 4.0, 3.0)], \
            "Balls stand-in (Balls.obj is missing from the reference checkout): island height field + 3 tessellated spheres, " \
            f"{s.n_triangles} triangles, Balls.mtl materials, 800x800, 4x4 rays/pixel, 1 light, shadows + reflection, max_lvl 3"
    if name == "balls_spheres":   # configs[1] read literally: "Balls.obj + Sphere primitives"
        s = scenes.balls_with_sphere_primitives()
        return s, 800, 800, 4, 3, (0.0, 2.6, 5.2), (0.0, 0.55, 0.0), [(2.5, 4.0, 3.0)], \
            f"Balls stand-in terrain ({s.n_triangles} triangles) + 3 analytic Sphere primitives (own semantics: the reference's Sphere.h is " \
            "orphaned), 800x800, 4x4 rays/pixel, 1 light, shadows + reflection, max_lvl 3"
    if name == "dodge":      # configs[2]
        s = host.Scene.load(os.path.join(ROOT, "tests", "golden", "scenes", "dodge.npz"))
        return s, 1920, 1080, 4, 10, (.75, .55, 1.1), (.07, 0, .23), None, \
            f"dodgeColorTest.obj ({s.n_triangles} triangles) 1920x1080, 4x4 rays/pixel, light at the eye, max_lvl 10"
    if name == "cube":       # configs[0]
        s = host.Scene.load(os.path.join(ROOT, "tests", "golden", "scenes", "cube.npz"))
        return s, 800, 800, 1, 10, (2.6, 2.4, 3.0), (.5, .5, .5), None, "cube.obj 800x800, 1 ray/pixel, light at the eye"
    if name == "sphere1m":   # configs[3]
        s = scenes.tessellated_sphere()
        return s, 3840, 2160, 4, 3, (0.0, 0.6, 3.4), (0, 0, 0), [(2.5, 4.0, 3.0)], \
            f"synthetic tessellated sphere ({s.n_triangles} triangles) 3840x2160, 4x4 rays/pixel, max_lvl 3"
    raise SystemExit(f"unknown workload {name}")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        busy = [x for x in sm if x > 0.5 * max(sm)] or sm
        reasons = [n for i, n in [(4, "hw_slowdown"), (5, "hw_thermal_slowdown"), (6, "sw_thermal_slowdown"), (7, "sw_power_cap")]
                   if any(r[i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[3]) for r in self.rows if r[3].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(self.rows[0][2]), "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def cpu_oracle():
    from oracle import pyoracle
    if os.path.exists(pyoracle.REF_SO):
        return pyoracle.RefOracle(), "reference"
    return pyoracle.PortOracle(), "port"


def lattice_pixels(W, H, k):
    return len(range(k // 2, H, k)) * len(range(k // 2, W, k))


def cpu_sample(scene, cam, pf, lvl, lights, k, threads, count=True):
    """Times the CPU reference on the pixel lattice (every k-th pixel of every k-th row, all pf*pf sub-samples of
    each) of the same frame; returns (seconds, rays, pixels, kind).  Rays are counted (untimed) by the port,
    which tests/test_oracle_golden.py pins bit-for-bit to the reference build."""
    from oracle import pyoracle
    O, kind = cpu_oracle()
    O.set_scene(scene)
    O.configure(cam.eye, lights, 63, lvl)
    t = time.perf_counter()
    O.render(cam.corners, cam.W, cam.H, pf, pf, y0=k // 2, ystep=k, x0=k // 2, xstep=k, threads=threads)
    dt = time.perf_counter() - t
    if not count:
        return dt, 0, lattice_pixels(cam.W, cam.H, k), kind
    P = pyoracle.PortOracle()
    P.set_scene(scene); P.configure(cam.eye, lights, 63, lvl); P.reset_counts()
    P.render(cam.corners, cam.W, cam.H, pf, pf, y0=k // 2, ystep=k, x0=k // 2, xstep=k, threads=threads)
    return dt, sum(P.ray_counts()), lattice_pixels(cam.W, cam.H, k), kind


def lattice_for_seconds(scene, cam, pf, lvl, lights, threads, seconds):
    """Pick the lattice pitch k so that one timed sample takes about `seconds` (calibrated on a coarse lattice)."""
    k0 = max(1, int(round((cam.W * cam.H / 512.0) ** 0.5)))
    t0, _, n0, _ = cpu_sample(scene, cam, pf, lvl, lights, k0, threads, count=False)
    target = max(64.0, n0 * seconds / max(t0, 1e-4))
    return max(1, int(np.ceil((cam.W * cam.H / target) ** 0.5)))


def run_reference(args, name):
    """--impl reference: the reference's CPU path on the host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from raytracert_b200 import host
    scene, W, H, pf, lvl, eye, center, lights, desc = workload(name)
    cam = host.Camera(W, H, eye, center)
    lights = [cam.eye] if lights is None else lights
    threads = os.cpu_count() or 1
    # each step costs a timed sample + the untimed ray count: keep the whole run near 2.5 minutes
    k = lattice_for_seconds(scene, cam, pf, lvl, lights, threads, 150.0 / (args.steps + args.warmup + 1))
    times, rays, npix, kind = [], 0, 0, "port"
    for i in range(args.warmup + args.steps):
        dt, r, npix, kind = cpu_sample(scene, cam, pf, lvl, lights, k, threads, count=(i == 0))   # same lattice, same rays every step
        rays = max(rays, r)
        if i >= args.warmup:
            times.append(dt)
    sec = float(np.mean(times))
    mrays = rays / sec / 1e6
    sample = (f"every {k}th pixel of every {k}th row of the same frame per step ({npix} of {W * H} pixels, all {pf * pf} sub-samples each), "
              f"{threads} OpenMP threads over pixels in the harness; ms_per_step is the sample's time scaled by {W * H}/{npix}")
    line = {"impl": "reference", "metric": METRIC, "value": mrays, "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3 * W * H / npix,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "name": name},
            "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="balls")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-accelerated", action="store_true", help="skip the separately reported tile-culling frames (profiling runs)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    args = ap.parse_args()
    guard_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args, args.workload)

    import torch
    from raytracert_b200 import binding, dist, host
    R, rank, world = dist.make_renderer()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if world == 1 and args.gpus > 1:
        raise SystemExit("for --gpus N > 1 launch with torchrun (one rank per GPU)")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    scene, W, H, pf, lvl, eye, center, lights, desc = workload(args.workload)
    cam = host.Camera(W, H, eye, center)
    lights = [cam.eye] if lights is None else lights
    R.upload_scene(scene)
    prm = binding.make_params(cam.corners, W, H, pf, pf, lvl, binding.RT_ALL_FEATURES, cam.eye, lights)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")   # > 126 MB L2

    def barrier():
        if world > 1:
            import torch.distributed as td
            td.barrier()
        torch.cuda.synchronize()
        R.sync()

    def frame_ms(e2e=False):
        """One timed frame.  Device-resident: events on the library's stream around rt_render.  e2e: host buffers in,
        host framebuffer out, through the public calls (rt_upload_scene + rt_render + rt_download_framebuffer)."""
        flush.zero_(); torch.cuda.synchronize()       # L2 flush between timed iterations (outside the timed region)
        if not e2e:
            R.event_record(0); R.render(prm, sync=False); R.event_record(1); R.sync()
            return R.event_elapsed_ms(0, 1)
        barrier()                                     # ranks start the step together (the all-gather would otherwise absorb their skew)
        t = time.perf_counter()
        R.upload_scene(scene); R.render(prm); R.download_into(fb_host)
        return (time.perf_counter() - t) * 1e3

    fb_pinned = torch.zeros((H, W, 3), dtype=torch.float32).pin_memory()   # the D2H target of the e2e steps is pinned host memory
    fb_host = fb_pinned.numpy()
    for _ in range(args.warmup):
        frame_ms()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t_wall = time.perf_counter()
    per_step = [frame_ms() for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.summary()
    st = R.stats()
    fp32_probe = R.probe_fp32_peak()   # measured packed-FMA ceiling of this device, for context beside the nominal peak
    # end-to-end through the public API with host buffers (scene H2D + frame + framebuffer D2H every step)
    e2e_steps = [frame_ms(e2e=True) for _ in range(max(3, min(args.steps, 5)))]
    fb_brute = fb_host.copy()
    # opt-in extra (not the contract path): conservative tile culling, same image bit for bit, reported separately
    cull_steps, cull_identical = [float("nan")], None
    if not args.no_accelerated:
        R.set_option(binding.RT_OPT_TILE_CULLING, 1)
        R.upload_scene(scene)                         # with the option set, the upload also sorts the tiles spatially
        for _ in range(3):
            frame_ms()
        barrier()
        cull_steps = [frame_ms() for _ in range(max(3, min(args.steps, 10)))]
        barrier()
        R.download_into(fb_host)
        cull_identical = bool(np.array_equal(fb_host.view(np.uint32), fb_brute.view(np.uint32)))
        R.set_option(binding.RT_OPT_TILE_CULLING, 0)
        R.upload_scene(scene)

    ms_dev = float(np.mean(per_step))
    ms_e2e = float(np.mean(e2e_steps))
    ms_cull = float(np.mean(cull_steps))
    counts = np.array([st["primary_rays"], st["shadow_rays"], st["bounce_rays"], st["exact_evals"]], np.float64)
    kinds = np.array([st["ms_trace"], st["ms_shadow"], st["ms_shade"], st["ms_resolve"], st["ms_gather"], st["ms_trace_primary"]], np.float64)
    if world > 1:
        import torch.distributed as td
        t = torch.tensor([ms_dev, ms_e2e, ms_cull] + list(kinds), dtype=torch.float64, device=f"cuda:{local}")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        ms_dev, ms_e2e, ms_cull, kinds = float(t[0]), float(t[1]), float(t[2]), t[3:].cpu().numpy()
        c = torch.tensor(counts, dtype=torch.float64, device=f"cuda:{local}")
        td.all_reduce(c, op=td.ReduceOp.SUM)
        counts = c.cpu().numpy()
    rays = float(counts[:3].sum())
    ntri = scene.n_triangles

    if rank == 0:
        peaks, peak_src = measured_peaks()
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        fp32_peak = sms * 128 * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12           # TFLOP/s per GPU at max clock
        # Scan kernels of the last frame.  Algorithmic flops = 42 per (ray, triangle) test (the minimal ray-dependent
        # form of rayIntersectTriangle for a GENERAL ray) x rays x triangles; every kernel performs every test of the reference.
        #   k_trace generic  : bounce levels (and the primary level when the pencil filter does not apply): 27 executed flop / test
        #   k_trace pencil   : primary rays, common-point filter in a projective chart: hot loop 6 FFMA2 per ray pair = 12 executed flop / test
        #   k_shadow         : any-hit; the reference's shadow rays are full scans, so its tests count in full although rays exit early
        variant = int(st["variant"])
        pencil_primary, pencil_shadow = bool(variant & 2), bool(variant & 4)
        ms_primary = float(kinds[5])
        ms_bounce = float(kinds[0] - kinds[5])

        def kernel_row(name, rays_gpu, ms, executed):
            a = FLOPS_PER_TEST * rays_gpu * ntri / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
            return {"kernel": name, "ms": ms, "tests_per_s": rays_gpu * ntri / (ms * 1e-3) if ms > 0 else 0.0, "achieved": a, "frac": a / fp32_peak,
                    "executed_flops_per_test": executed, "executed_frac": a * executed / FLOPS_PER_TEST / fp32_peak}
        rows = [kernel_row("k_trace primary (%s filter)" % ("pencil" if pencil_primary else "generic"), counts[0] / world, ms_primary, 12 if pencil_primary else 27),
                kernel_row("k_trace bounce levels (generic filter)", counts[2] / world, ms_bounce, 27),
                kernel_row("k_shadow any-hit (%s filter)" % ("pencil" if pencil_shadow else "generic"), counts[1] / world, float(kinds[1]), 12 if pencil_shadow else 27)]
        dom = max(rows, key=lambda r: r["ms"])
        ach = dom["achieved"]
        roof = {"bound": "fp32", "kernel": dom["kernel"] + " -- the launch kind with the largest share of the frame", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": ach / fp32_peak,
                # rt_probe_fp32_peak(): what a register-resident FFMA/FFMA2 loop sustains on this device right now (TFLOP/s)
                "peak_fma_loop_measured": fp32_probe, "frac_of_measured_fma_loop": ach / fp32_probe if fp32_probe > 0 else None,
                # dram__bytes_read.sum + dram__bytes_write.sum of the largest launch of that kind (level-1 bounce scan of chunk 0,
                # 4.34 M rays, 134.6 ms), one `ncu --set full` capture of this command (profiles/r1h_k_trace_bounce_full.txt);
                # only meaningful for the default workload
                "traffic": 182.7e6 if args.workload == "balls" and world == 1 else None, "traffic_unit": "B per launch (level-1 bounce scan, chunk 0)",
                "peak_source": f"148 SMs x 128 lanes x 2 x sm_max_mhz ({peak_src} MEASURED_PEAKS.json clock)",
                "frac_at_measured_clock": (ach / (fp32_peak * clocks["sm_mhz"] / clocks["sm_max_mhz"])) if clocks.get("sm_mhz") else None,
                "executed_flops_per_test": dom["executed_flops_per_test"], "executed_frac": dom["executed_frac"],
                "frame_achieved": FLOPS_PER_TEST * rays * ntri / world / (ms_dev * 1e-3) / 1e12,
                "by_kernel": rows,
                "note": "achieved/frac count 42 algorithmic flop per test (general-ray form, SURVEY 8d). The pencil kernels do the same tests "
                        "with 12 executed flop in the hot loop (rays through a common point need no origin arithmetic; the distance clause runs in the cold path), so their algorithmic rate -- and "
                        "frame_achieved -- can exceed the FP32 peak; executed_frac is the FMA-pipe-level figure for every row.",
                "ms_by_kernel": dict(zip(["k_trace", "k_shadow", "k_shade", "k_resolve", "gather", "k_trace_primary"], [float(x) for x in kinds]))}
        line = {"metric": METRIC, "value": rays / ms_dev / 1e3, "unit": "Mrays/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": desc, "name": args.workload, "triangles": ntri, "rays_per_frame": rays,
                           "primary": counts[0], "shadow": counts[1], "bounce": counts[2], "exact_reevaluations": counts[3],
                           "primary_mrays_per_s": counts[0] / ms_dev / 1e3, "parallelism": f"rows interleaved over {world} GPU(s)",
                           "filter": {"primary": "pencil" if variant & 2 else "generic", "shadow": "pencil" if variant & 4 else "generic", "bounce": "generic",
                                      "clause_free": bool(variant & 1)},
                           "l2": "flushed between timed iterations (256 MiB write)", "wall_s_timed_region": t_wall},
                "clocks": clocks,
                "e2e": {"value": rays / ms_e2e / 1e3, "unit": "Mrays/s", "ms_per_step": ms_e2e, "ms_steps_rank0": [round(x, 2) for x in e2e_steps],
                        "h2d_bytes_per_step": int(ntri * (4 * 16 + 4) + scene.materials.shape[0] * 64 + 432), "d2h_bytes_per_step": int(fb_host.nbytes),
                        "path": "rt_upload_scene + rt_render + rt_download_framebuffer with host buffers, every step"},
                "gpu_launches": int(st["n_launches"]) * args.steps,
                "accelerated": {"option": "RT_OPT_TILE_CULLING (opt-in; not the brute-force contract path, not used for value/e2e/roofline)",
                                "ms_per_step": ms_cull, "value": rays / ms_cull / 1e3, "unit": "Mrays/s", "image_bit_identical_to_brute_force": cull_identical},
                "roofline": roof}
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            k = lattice_for_seconds(scene, cam, pf, lvl, lights, threads, args.cpu_seconds)
            dt, r, npix, kind = cpu_sample(scene, cam, pf, lvl, lights, k, threads)
            line["cpu_baseline"] = {"value": r / dt / 1e6, "unit": "Mrays/s", "cores": threads, "kind": kind,
                                    "sample": f"every {k}th pixel of every {k}th row of the same frame ({npix} of {W * H} pixels, all {pf * pf} sub-samples "
                                              f"each), {dt:.1f} s; the reference's own raytracing.cpp/mesh.cpp (-O2 -ffp-contract=off), OpenMP over pixels in the harness",
                                    "ms_per_frame_extrapolated": dt * 1e3 * W * H / npix}
        emit(line)
    R.shutdown()
    if world > 1:
        import torch.distributed as td
        td.barrier()
        td.destroy_process_group()


if __name__ == "__main__":
    main()
