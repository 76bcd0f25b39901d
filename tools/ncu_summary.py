"""Print the metrics we care about from an .ncu-rep (raw page), one block per captured launch.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [extra-substring ...]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
extra = sys.argv[2:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
EXACT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
         "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
         "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active",
         "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed_per_inst_executed.pct",
         "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
         "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
         "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
         "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
         "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
         "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("=" * 100)
    for i, h in enumerate(hdr):
        if h in EXACT or any(e in h for e in extra) or ("warp_issue_stalled" in h and h.endswith("_per_warp_active.pct")):
            try:
                v = float(r[i].replace(",", ""))
                if "stalled" in h and v < 1.0:
                    continue
                print(f"{h:100s} {v:>18,.3f} {units[i]}")
            except ValueError:
                print(f"{h:100s} {r[i][:70]}")
