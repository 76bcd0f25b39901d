"""Collects the measurement matrix (tools/gpu/matrix.sh, gpurun_out/<tag>_n{1,2,4,8}/) into profiles/r2_matrix/ and prints the
markdown tables of DESIGN.md section 6.    python tools/matrix_table.py r2m gpurun_out"""
import json, os, shutil, sys

tag, src = (sys.argv + ["r2m", "gpurun_out"])[1:3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "profiles", "r2_matrix")
os.makedirs(out, exist_ok=True)


def load(path):
    s = open(path).read()
    return json.loads(s[s.index("{"):]) if "{" in s else None


rows_c = {}
for n in (1, 2, 4, 8):
    d = os.path.join(src, f"{tag}_n{n}")
    if not os.path.isdir(d):
        continue
    for f in sorted(os.listdir(d)):
        if f.endswith(".json") or f.endswith(".jsonl") or f in ("summary.txt", "nvidia_smi.csv", "pytest_multi_gpu.log"):
            dst = f if f"_n{n}" in f else f"n{n}_{f}"
            shutil.copy(os.path.join(d, f), os.path.join(out, dst))
    rows_c[n] = d

print("| config | GPUs | ms/frame (device) | Mrays/s | e2e ms (upload / render / download) | FP32 at the pipe (executed) | algorithmic ratio | parity vs reference pin | CPU reference, same box |")
print("|---|---|---|---|---|---|---|---|---|")
for cfg, f in (("C1 cube.obj 800² 1 spp", "C1_bench_n{n}.json"), ("C2 Balls stand-in 800² 16 spp", "C2_bench_n{n}.json"),
               ("C2, one process drives all GPUs", "C2_bench_single_process_n{n}.json"), ("C3 dodgeColorTest 1920×1080 16 spp", "C3_bench_n{n}.json")):
    for n, d in rows_c.items():
        p = os.path.join(d, f.format(n=n))
        if not os.path.exists(p):
            continue
        b = load(p)
        if not b:
            continue
        e = b["e2e"]; r = b["roofline"]; par = b.get("parity") or {}
        ex = r.get("frame_executed_frac")
        if ex is None:
            ex = sum(k.get("executed_frac_from_hot_loop", k.get("frac", 0.0)) * k["ms"] for k in r["by_kernel"]) / max(sum(k["ms"] for k in r["by_kernel"]), 1e-9)
        ptxt = "—" if not par.get("against") else f"{par['rows_checked']} rows: {par['id_mismatches']} id mismatches, {par['u8_off_by_more_than_1']} u8 off by > 1"
        cb = b.get("cpu_baseline")
        ctxt = "—" if not cb else f"{cb['value']:.4g} Mrays/s on {cb['cores']} cores ({cb['one_thread']['value']:.3g} on 1 thread)"
        print(f"| {cfg} | {n} | {b['ms_per_step']:.3f} | {b['value']:.1f} | {e['ms_per_step']:.2f} ({e['upload_ms']:.2f} / {e['render_ms']:.2f} / {e['download_ms']:.2f}) | "
              f"{ex:.3f} | {r.get('algorithmic_ratio', r['frac']):.2f} ({r['kernel'][:22]}…) | {ptxt} | {ctxt} |")

print()
print("| C4 1 M-triangle sphere 3840×2160 16 spp | GPUs | brute force ms | Mrays/s | executed FP32 frac | primary / bounce / shadow ms | tile culling ms | culled == brute bits | oracle lattice | CPU reference |")
print("|---|---|---|---|---|---|---|---|---|---|")
for n, d in rows_c.items():
    p = os.path.join(d, f"C4_sphere1m_n{n}.json")
    if not os.path.exists(p):
        continue
    b = load(p)
    if not b:
        continue
    bf, tc, ol = b["brute_force"], b.get("tile_culling", {}), b.get("oracle_lattice", {})
    cb = b.get("cpu_baseline")
    ctxt = "—" if not cb else f"{cb['value']:.3g} Mrays/s on {cb['cores']} cores ({cb['one_thread']['value']:.3g} on 1 thread); {cb['ms_per_frame_extrapolated'] / 3.6e6:.0f} h per frame"
    print(f"| | {n} | {bf['ms_per_frame']:.0f} | {bf['Mrays_per_s']:.2f} | {bf['executed_frac_of_fp32_peak']:.3f} | {' / '.join(f'{x:.0f}' for x in bf['ms_primary_bounce_shadow'])} | "
          f"{tc.get('ms_per_frame', float('nan')):.0f} | {b.get('culling_bit_identical')} | {ol.get('pixels')} px: max |ΔRGB| {ol.get('max_abs_rgb_diff', 0):.1e}"
          f"{', ' + str(ol['id_mismatches']) + ' id mismatches' if 'id_mismatches' in ol else ''} | {ctxt} |")

print()
table = {}
for n, d in rows_c.items():
    p = os.path.join(d, f"C5_n{n}.jsonl")
    if not os.path.exists(p):
        continue
    for l in open(p):
        if l.startswith("{"):
            r = json.loads(l)
            table.setdefault((r["size"], r["spp"]), {})[n] = r
ns = sorted(rows_c)
print("| C5: size | spp | " + " | ".join(f"N={n}: ms (Mrays/s, exec. frac) [culled ms]" for n in ns) + " |")
print("|---|---|" + "---|" * len(ns))
for (size, spp) in sorted(table):
    cells = []
    for n in ns:
        r = table[(size, spp)].get(n)
        if not r:
            cells.append("—")
            continue
        b = r["brute_force"]; c = r.get("tile_culling")
        cells.append(f"{b['ms_per_frame']:.1f} ({b['Mrays_per_s']:.0f}, {b['executed_frac_of_fp32_peak']:.2f})" + (f" [{c['ms_per_frame']:.1f}]" if c else ""))
    print(f"| {size}² | {spp} | " + " | ".join(cells) + " |")
