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


def lattice_pitch(W, H, pixels=1600.0):
    """One fixed pitch per workload for every CPU timing of it (cpu_baseline at any N, --impl reference): about 1600
    pixels of the frame (every 20th pixel of every 20th row for 800x800)."""
    return max(1, int(round((W * H / pixels) ** 0.5)))


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


def config_for(name, desc, scene, W, H, pf, lvl, n_lights, n_gpus):
    """`config` of the JSON line -- static description of the workload only, identical in both arms."""
    return {"workload": desc, "name": name, "triangles": int(scene.n_triangles), "width": W, "height": H, "rays_per_pixel": pf * pf,
            "max_lvl": lvl, "lights": n_lights, "features": "ambient+diffuse+specular+reflection+shadows+refraction (all toggles on)",
            "parallelism": f"rows interleaved over {n_gpus} GPU(s), scene replicated"}


def cpu_baseline_block(scene, cam, pf, lvl, lights, threads, pixels=1600.0):
    """The reference's CPU path on this box: all host threads on the fixed lattice, plus a 1-thread figure on a 16x
    sparser lattice of the same frame (SURVEY 8d asks for both)."""
    W, H = cam.W, cam.H
    k = lattice_pitch(W, H, pixels)
    dt, r, npix, kind = cpu_sample(scene, cam, pf, lvl, lights, k, threads)
    k1 = 4 * k
    dt1, r1, npix1, _ = cpu_sample(scene, cam, pf, lvl, lights, k1, 1)
    what = "the reference's own raytracing.cpp/mesh.cpp (-O2 -ffp-contract=off), OpenMP over pixels in the harness" if kind == "reference" \
        else "plain-C port of the reference (oracle/rt_oracle.c), OpenMP over pixels"
    return {"value": r / dt / 1e6, "unit": "Mrays/s", "cores": threads, "kind": kind,
            "sample": f"every {k}th pixel of every {k}th row of the same frame ({npix} of {W * H} pixels, all {pf * pf} sub-samples each), {dt:.1f} s; {what}",
            "ms_per_frame_extrapolated": dt * 1e3 * W * H / npix,
            "one_thread": {"value": r1 / dt1 / 1e6, "unit": "Mrays/s", "cores": 1,
                           "sample": f"every {k1}th pixel of every {k1}th row ({npix1} pixels), {dt1:.1f} s"}}


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
    k = lattice_pitch(W, H)   # the same lattice as the product arm's cpu_baseline
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
    k1 = 4 * k
    dt1, r1, npix1, _ = cpu_sample(scene, cam, pf, lvl, lights, k1, 1)
    line = {"impl": "reference", "metric": METRIC, "value": mrays, "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3 * W * H / npix,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_for(name, desc, scene, W, H, pf, lvl, len(lights), args.gpus),
            "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": threads, "kind": kind, "sample": sample,
                             "one_thread": {"value": r1 / dt1 / 1e6, "unit": "Mrays/s", "cores": 1,
                                            "sample": f"every {k1}th pixel of every {k1}th row ({npix1} pixels), {dt1:.1f} s"}},
            "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def pin_path(name):
    return os.path.join(ROOT, "tests", "golden", "pins", name + ".npz")


def parity_block(R, name, prm_ids, rank, world, single_process):
    """Compares one more frame (rendered after the timed region, with per-sample ids kept) with the frame the UNMODIFIED
    reference produced for this workload (tests/golden/pins/<name>.npz, tools/make_headline_pin.py).  Ids: every rank
    checks its own rows (per-row CRC, and sample by sample), the counts are summed over the ranks.  u8 image: rank 0, on
    the frame as the all-gather + de-interleave assembled it.  Returns the block (rank 0) or None."""
    import zlib
    path = pin_path(name)
    if not os.path.exists(path):
        return {"against": None, "note": f"no pin for workload {name} (tests/golden/pins)"}
    z = np.load(path)
    rows = z["rows"].astype(np.int64)
    H, W = int(z["H"]), int(z["W"])
    R.render(prm_ids)
    rgb, prim = R.download(want_prim_id=True)
    prim = prim.reshape(H, -1)
    mine = rows if (single_process or world == 1) else rows[rows % world == rank]
    idx = {int(y): i for i, y in enumerate(rows)}
    ids = np.frombuffer(zlib.decompress(z["ids_z"].tobytes()), "<i4").reshape(len(rows), -1) if "ids_z" in z.files else None
    bad_rows = bad_ids = 0
    for y in mine:
        if np.uint32(zlib.crc32(prim[y].astype("<i4").tobytes())) != z["id_crc"][idx[int(y)]]:
            bad_rows += 1
            if ids is not None:
                bad_ids += int(np.count_nonzero(ids[idx[int(y)]] != prim[y]))
    counts = np.array([len(mine), bad_rows, bad_ids], np.float64)
    if world > 1 and not single_process:
        import torch
        import torch.distributed as td
        t = torch.tensor(counts, dtype=torch.float64, device="cuda" if td.get_backend() == "nccl" else "cpu")
        td.all_reduce(t, op=td.ReduceOp.SUM)
        counts = t.cpu().numpy()
    if rank != 0:
        return None
    u8 = R.download_u8()
    d = np.abs(u8[rows].astype(np.int32) - z["u8"].astype(np.int32))
    return {"against": f"tests/golden/pins/{name}.npz -- the unmodified reference (oracle/_ref), {len(rows)} of the frame's {H} rows",
            "rows_checked": int(counts[0]), "rows_in_frame": H, "id_rows_mismatching": int(counts[1]), "id_mismatches": int(counts[2]) if ids is not None else None,
            "samples_checked": int(counts[0]) * prim.shape[1], "u8_off_by_more_than_1": int(np.count_nonzero(d > 1)),
            "u8_pixels_differing_frac": float(np.mean(np.any(d > 0, axis=2))), "float_rgb_max_abs_diff_note": "colour differs from the reference only through CUDA powf vs glibc powf (<= 2e-7)",
            "n_gpus": world}


def kernel_counters(name):
    """Per-kernel ncu counters of this workload (profiles/r2_kernel_counters.json, written from ncu captures by
    tools/ncu_counters.py): measured, but NOT in this run."""
    p = os.path.join(ROOT, "profiles", "r2_kernel_counters.json")
    if not os.path.exists(p):
        return {}, None
    return json.load(open(p)).get(name, {}), "profiles/r2_kernel_counters.json"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="balls")
    ap.add_argument("--single-process", action="store_true",
                    help="drive all --gpus N devices from THIS process (rt_init(N), ncclCommInitAll: the C++ drop-in's mode) instead of one torchrun rank per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-run comparison with the reference's pinned frame")
    ap.add_argument("--no-accelerated", action="store_true", help="skip the separately reported tile-culling frames (profiling runs)")
    args = ap.parse_args()
    guard_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args, args.workload)

    import torch
    from raytracert_b200 import binding, dist, host
    single = bool(args.single_process)
    if single:
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            raise SystemExit("--single-process is not a torchrun mode")
        R, rank, world = binding.Renderer(args.gpus), 0, args.gpus
    else:
        R, rank, world = dist.make_renderer()
        if world != args.gpus and world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
        if world == 1 and args.gpus > 1:
            raise SystemExit("for --gpus N > 1 launch with torchrun (one rank per GPU), or pass --single-process")
    local = 0 if single else int(os.environ.get("LOCAL_RANK", "0"))
    my_devices = list(range(world)) if single else [local]
    torch.cuda.set_device(local)
    scene, W, H, pf, lvl, eye, center, lights, desc = workload(args.workload)
    cam = host.Camera(W, H, eye, center)
    lights = [cam.eye] if lights is None else lights
    R.upload_scene(scene)
    prm = binding.make_params(cam.corners, W, H, pf, pf, lvl, binding.RT_ALL_FEATURES, cam.eye, lights)
    flush = [torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{i}") for i in my_devices]   # > 126 MB L2, one per driven device
    multi_rank = world > 1 and not single

    def barrier():
        if multi_rank:
            import torch.distributed as td
            td.barrier()
        for i in my_devices:
            torch.cuda.synchronize(i)
        R.sync()

    e2e_parts = {"upload_ms": [], "render_ms": [], "download_ms": []}

    def frame_ms(e2e=False):
        """One timed frame.  Device-resident: events on the library's stream around rt_render.  e2e: host buffers in,
        host framebuffer out, through the public calls (rt_upload_scene + rt_render + rt_download_framebuffer)."""
        for f in flush:                               # L2 flush between timed iterations (outside the timed region)
            f.zero_()
        for i in my_devices:
            torch.cuda.synchronize(i)
        if not e2e:
            R.event_record(0); R.render(prm, sync=False); R.event_record(1); R.sync()
            return R.event_elapsed_ms(0, 1)
        barrier()                                     # ranks start the step together (the all-gather would otherwise absorb their skew)
        t0 = time.perf_counter()
        R.upload_scene(scene)
        t1 = time.perf_counter()
        R.render(prm)
        t2 = time.perf_counter()
        if rank == 0:                                 # the frame is the job's result: one host reads it back (every rank holds it after the all-gather)
            R.download_into(fb_host)
        t3 = time.perf_counter()
        for k, v in zip(("upload_ms", "render_ms", "download_ms"), (t1 - t0, t2 - t1, t3 - t2)):
            e2e_parts[k].append(v * 1e3)
        return (t3 - t0) * 1e3

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
    frame_ms(e2e=True)                 # one untimed e2e step (first use of the staging buffers)
    for v in e2e_parts.values():
        v.clear()
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
        if rank == 0:
            R.download_into(fb_host)
        cull_identical = bool(np.array_equal(fb_host.view(np.uint32), fb_brute.view(np.uint32)))
        R.set_option(binding.RT_OPT_TILE_CULLING, 0)
        R.upload_scene(scene)
    # parity against the reference's own frame, at every N (after the timed region; ids kept for this one frame)
    parity = None
    if not args.no_parity:
        prm_ids = binding.make_params(cam.corners, W, H, pf, pf, lvl, binding.RT_ALL_FEATURES, cam.eye, lights, want_prim_id=True)
        parity = parity_block(R, args.workload, prm_ids, rank, world, single)

    ms_dev = float(np.mean(per_step))
    ms_e2e = float(np.mean(e2e_steps))
    ms_cull = float(np.mean(cull_steps))
    parts = [float(np.mean(e2e_parts[k])) for k in ("upload_ms", "render_ms", "download_ms")]
    counts = np.array([st["primary_rays"], st["shadow_rays"], st["bounce_rays"], st["exact_evals"], st["mirror_rays"], st["thread_pencil_rays"]], np.float64)
    kinds = np.array([st["ms_trace"], st["ms_shadow"], st["ms_shade"], st["ms_resolve"], st["ms_gather"], st["ms_trace_primary"], st["ms_trace_mirror"],
                      st["ms_trace_thread"]], np.float64)
    if multi_rank:
        import torch.distributed as td
        t = torch.tensor([ms_dev, ms_e2e, ms_cull] + parts + list(kinds), dtype=torch.float64, device=f"cuda:{local}")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        ms_dev, ms_e2e, ms_cull, parts, kinds = float(t[0]), float(t[1]), float(t[2]), [float(x) for x in t[3:6]], t[6:].cpu().numpy()
        c = torch.tensor(counts, dtype=torch.float64, device=f"cuda:{local}")
        td.all_reduce(c, op=td.ReduceOp.SUM)
        counts = c.cpu().numpy()
    rays = float(counts[:3].sum())
    ntri = scene.n_triangles

    if rank == 0:
        peaks, peak_src = measured_peaks()
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        fp32_peak = sms * 128 * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12           # TFLOP/s per GPU at max clock
        ncu, ncu_src = kernel_counters(args.workload)
        # Scan kernels of the last frame.  Algorithmic flops = 42 per (ray, triangle) test (the minimal ray-dependent
        # form of rayIntersectTriangle for a GENERAL ray) x rays x triangles; every kernel performs every test of the reference.
        #   k_trace generic  : bounce levels (and the primary level when the pencil filter does not apply): 27 executed flop / test
        #   k_trace pencil   : primary rays, common-point filter in a projective chart: hot loop 6 FFMA2 per ray pair = 12 executed flop / test
        #   k_shadow         : any-hit; the reference's shadow rays are full scans, so its tests count in full although rays exit early
        variant = int(st["variant"])
        pencil_primary, pencil_shadow = bool(variant & 2), bool(variant & 4)
        ms_primary = float(kinds[5])
        ms_mirror = float(kinds[6])
        ms_thread = float(kinds[7])
        ms_bounce = float(kinds[0] - kinds[5] - kinds[6] - kinds[7])     # generic scans of the bounce levels

        def kernel_row(key, name, rays_gpu, ms, executed):
            alg = FLOPS_PER_TEST * rays_gpu * ntri / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
            ex = alg * executed / FLOPS_PER_TEST
            m = ncu.get(key, {})
            return {"kernel": name, "ms": ms, "tests_per_s": rays_gpu * ntri / (ms * 1e-3) if ms > 0 else 0.0,
                    "executed_flops_per_test": executed, "achieved": ex, "frac": ex / fp32_peak,          # executed at the FMA pipe
                    "algorithmic_achieved": alg, "algorithmic_ratio": alg / fp32_peak,                   # 42 flop per test; NOT a utilisation (can exceed 1)
                    "fma_pipe_active_ncu": m.get("fma_pipe_cycles_active_pct"), "issue_active_ncu": m.get("issue_active_pct"),
                    "dram_bytes_ncu": m.get("dram_bytes"), "ncu_launch": m.get("launch")}
        # executed flops per (ray, triangle) test in the hot loops (SASS: profiles/r2_sass_hot_loops.txt):
        #   generic filter 11 FFMA2 + 3 FMUL2 + 2 FADD2 per ray pair = 27; pencil filters 6 FFMA2 per ray pair = 12;
        #   thread pencils 9 FFMA per ray + (21 FFMA + 3 FADD) per triangle and thread of 8 rays = 23.6
        rows = [kernel_row("primary", "k_trace primary (%s filter)" % ("pencil" if pencil_primary else "generic"), counts[0] / world, ms_primary, 12 if pencil_primary else 27),
                kernel_row("bounce", "k_trace bounce levels (generic filter)", (counts[2] - counts[4] - counts[5]) / world, ms_bounce, 27),
                kernel_row("thread", "k_trace_tp level-1 rays grouped by reflector (thread pencils)", counts[5] / world, ms_thread, 23.6),
                kernel_row("mirror", "k_trace level-1 rays of plane groups (mirror pencil)", counts[4] / world, ms_mirror, 12),
                kernel_row("shadow", "k_shadow any-hit (%s filter)" % ("pencil" if pencil_shadow else "generic"), counts[1] / world, float(kinds[1]), 12 if pencil_shadow else 27)]
        dom = max(rows, key=lambda r: r["ms"])
        roof = {"bound": "fp32", "kernel": dom["kernel"] + " -- the launch kind with the largest share of the frame",
                # achieved = FP32 flops the kernel EXECUTES (hot-loop instruction mix x measured test rate) / its device time; frac = achieved / peak
                "achieved": dom["achieved"], "peak": fp32_peak, "unit": "TFLOP/s", "frac": dom["frac"],
                "executed_flops_per_test": dom["executed_flops_per_test"], "fma_pipe_active_ncu": dom["fma_pipe_active_ncu"],
                # the same launches counted with SURVEY 8d's 42 algorithmic flop per test (general-ray form): what the reference's arithmetic would need
                "algorithmic_achieved": dom["algorithmic_achieved"], "algorithmic_ratio": dom["algorithmic_ratio"],
                # rt_probe_fp32_peak(): what a register-resident FFMA/FFMA2 loop sustains on this device right now (TFLOP/s)
                "peak_fma_loop_measured": fp32_probe, "frac_of_measured_fma_loop": dom["achieved"] / fp32_probe if fp32_probe > 0 else None,
                # dram__bytes_read.sum + dram__bytes_write.sum of the largest launch of that kind, from an ncu capture of this workload
                # (not measured in this run); null when there is no capture for the workload / GPU count
                "traffic": dom["dram_bytes_ncu"] if world == 1 else None, "traffic_source": ncu_src, "traffic_launch": dom["ncu_launch"],
                "peak_source": f"{sms} SMs x 128 lanes x 2 x sm_max_mhz ({peak_src} MEASURED_PEAKS.json clock; no FP32 entry there, so the nominal figure at the measured max clock)",
                "frac_at_measured_clock": (dom["achieved"] / (fp32_peak * clocks["sm_mhz"] / clocks["sm_max_mhz"])) if clocks.get("sm_mhz") else None,
                "frame_executed_frac": sum(r["achieved"] * r["ms"] for r in rows) / max(sum(r["ms"] for r in rows), 1e-9) / fp32_peak,
                "by_kernel": rows,
                "note": "achieved/frac are EXECUTED FP32 flops at the FMA pipe (hot-loop instruction mix x measured test rate), fma_pipe_active_ncu is ncu's own pipe counter for "
                        "the largest launch of the kind. algorithmic_* count 42 flop per test (SURVEY 8d, general-ray form): every kernel performs every (ray, triangle) test of the "
                        "reference, the pencil kernels in 12 executed flop (rays through a common point need no origin arithmetic), so that quotient can exceed 1 and is not a utilisation.",
                "ms_by_kernel": dict(zip(["k_trace", "k_shadow", "k_shade", "k_resolve", "gather", "k_trace_primary", "k_trace_mirror", "k_trace_thread"], [float(x) for x in kinds]))}
        line = {"metric": METRIC, "value": rays / ms_dev / 1e3, "unit": "Mrays/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_for(args.workload, desc, scene, W, H, pf, lvl, len(lights), world),
                "run": {"mode": "single process, rt_init(N)" if single else "one process per GPU (torchrun), rt_init_rank", "rays_per_frame": rays,
                        "primary": counts[0], "shadow": counts[1], "bounce": counts[2], "bounce_served_by_mirror_pencils": counts[4],
                        "bounce_served_by_thread_pencils": counts[5], "exact_reevaluations": counts[3],
                        "primary_mrays_per_s": counts[0] / ms_dev / 1e3,
                        "filter": {"primary": "pencil" if variant & 2 else "generic", "shadow": "pencil" if variant & 4 else "generic",
                                   "bounce": "generic" + (" + mirror pencils" if variant & 32 else "") + (" + thread pencils" if variant & 64 else ""),
                                   "clause_free": bool(variant & 1), "pencil_without_premise": bool(variant & 8), "graph_replay": bool(variant & 16)},
                        "l2": "flushed between timed iterations (256 MiB write per device)", "wall_s_timed_region": t_wall},
                "clocks": clocks,
                "e2e": {"value": rays / ms_e2e / 1e3, "unit": "Mrays/s", "ms_per_step": ms_e2e, "ms_steps_rank0": [round(x, 2) for x in e2e_steps],
                        "upload_ms": parts[0], "render_ms": parts[1], "download_ms": parts[2],
                        "h2d_bytes_per_step": int(ntri * (4 * 16 + 4) + scene.materials.shape[0] * 64 + 432), "d2h_bytes_per_step": int(fb_host.nbytes),
                        "path": "rt_upload_scene (every rank: the scene is replicated) + rt_render + rt_download_framebuffer (rank 0: the job's one result) with host buffers, every step; max over ranks of each part"},
                "gpu_launches": int(st["n_launches"]) * args.steps,
                "accelerated": {"option": "RT_OPT_TILE_CULLING (opt-in; not the brute-force contract path, not used for value/e2e/roofline)",
                                "ms_per_step": ms_cull, "value": rays / ms_cull / 1e3, "unit": "Mrays/s", "image_bit_identical_to_brute_force": cull_identical},
                "parity": parity,
                "roofline": roof}
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_block(scene, cam, pf, lvl, lights, os.cpu_count() or 1)
        emit(line)
    R.shutdown()
    if multi_rank:
        import torch.distributed as td
        td.barrier()
        td.destroy_process_group()


if __name__ == "__main__":
    main()
