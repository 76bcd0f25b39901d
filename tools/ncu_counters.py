"""profiles/r2_kernel_counters.json from an ncu launch list with counters (long CSV: one row per launch and metric).

    ncu --metrics gpu__time_duration.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,... --clock-control none \
        -k regex:"k_trace|k_shadow" --csv --log-file gpurun_out/x/inst_counts.csv python tools/prof_one.py 4
    python tools/ncu_counters.py balls gpurun_out/x/inst_counts.csv [--skip-launches N]

For each scan launch kind of the workload (primary / bounce / shadow) it keeps the LARGEST launch: duration, FMA-pipe
cycles active, instruction counts, DRAM bytes.  bench.py reads the file for `roofline.traffic` and the per-kernel
`fma_pipe_active_ncu` (measured by ncu, not in the bench run -- the bench line says so)."""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r2_kernel_counters.json")


def kind_of(name):
    m = re.match(r"(?:void )?(?:rt::)?(k_trace_tp|k_trace|k_shadow)<([^>]*)>", name)
    if not m:
        return None
    args = [a.strip() for a in m.group(2).split(",")]
    if m.group(1) == "k_shadow":
        return "shadow"
    if m.group(1) == "k_trace_tp":
        return "thread"
    if args[3] in ("1", "true"):
        return "primary"
    return "mirror" if len(args) > 6 and args[6] in ("1", "true") else "bounce"   # queued rays through the pencil filter: a plane group's mirror pencil


def main():
    workload, path = sys.argv[1], sys.argv[2]
    skip = int(sys.argv[sys.argv.index("--skip-launches") + 1]) if "--skip-launches" in sys.argv else 0
    lines = open(path).read().splitlines()
    i0 = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    L = collections.OrderedDict()
    for r in csv.DictReader(lines[i0:]):
        d = L.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        try:
            d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            pass
    best = {}
    for k, d in L.items():
        kind = kind_of(d["name"])
        if kind is None or k < skip:
            continue
        if kind not in best or d["gpu__time_duration.sum"] > best[kind]["gpu__time_duration.sum"]:
            best[kind] = dict(d, launch_id=k)
    out = json.load(open(OUT)) if os.path.exists(OUT) else {}
    w = out.setdefault(workload, {})
    for kind, d in best.items():
        g = lambda n: d.get(n)
        rd, wr = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
        w[kind] = {"kernel": d["name"], "launch": f"launch {d['launch_id']} of {os.path.basename(path)}: the largest {kind} launch, {d['gpu__time_duration.sum'] / 1e6:.2f} ms under ncu",
                   "ms_under_ncu": d["gpu__time_duration.sum"] / 1e6,
                   "fma_pipe_cycles_active_pct": g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                   "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   "inst_executed": g("smsp__inst_executed.sum"), "inst_pipe_fma": g("sm__inst_executed_pipe_fma.sum"),
                   "inst_pipe_alu": g("sm__inst_executed_pipe_alu.sum"), "inst_pipe_xu": g("sm__inst_executed_pipe_xu.sum"),
                   "inst_pipe_lsu": g("sm__inst_executed_pipe_lsu.sum"),
                   "scalar_ffma_fmul_fadd_thread_insts": [g("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"), g("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum"),
                                                          g("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum")],
                   "dram_bytes": (rd + wr) if rd is not None and wr is not None else None, "dram_read_bytes": rd, "dram_write_bytes": wr,
                   "source": os.path.relpath(path, ROOT)}
    json.dump(out, open(OUT, "w"), indent=1)
    for kind, v in w.items():
        print(kind, v["ms_under_ncu"], v["fma_pipe_cycles_active_pct"], v["dram_bytes"])


if __name__ == "__main__":
    main()
