"""Writes a text summary of an .ncu-rep (key raw metrics per kernel + per-source-line table of the first kernel).
usage: python tools/ncu_summary.py report.ncu-rep > profiles/xxx.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                     text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_op_global_red.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_op_red.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_atom.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
print("# ncu summary of %s (ncu --set full --clock-control none; one launch per kernel; cold-cache replay)" % rep.split("/")[-1])
for r in rows[2:]:
    print()
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print("%-70s %s %s" % (w, r[i][:110], units[i]))
print()
print(subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_lines.py"), rep, "30"], stdout=subprocess.PIPE,
                     text=True).stdout)
