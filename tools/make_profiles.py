"""Turns the artefacts tools/collect_profiles.sh left in gpurun_out/ into the tracked summaries under profiles/."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def summary(rep, out):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(G, rep)],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    open(os.path.join(P, out), "w").write(txt)


def raw_metrics(rep):
    raw = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}

    def get(name):
        i = hdr.index(name)
        return float(r[i].replace(",", "")) * mult.get(units[i], 1)
    return get


summary("r01_window.ncu-rep", "r01_std_grid_window_f32_continuum.txt")
summary("r01_iw.ncu-rep", "r01_imaging_weight_kernels.txt")
summary("r01_fused.ncu-rep", "r01_std_grid_window_fused_image_psf_f32.txt")
summary("r01_dr.ncu-rep", "r01_direction_rotate_phasor_f32.txt")
shutil.copy(os.path.join(G, "r01_launches.csv"), os.path.join(P, "r01_launches_bench_steps2.csv"))
shutil.copy(os.path.join(G, "r01_next_launches.csv"), os.path.join(P, "r01_gcf_launches.csv"))
if os.path.exists(os.path.join(G, "r01_apply_flags_launches.csv")):
    shutil.copy(os.path.join(G, "r01_apply_flags_launches.csv"), os.path.join(P, "r01_apply_flags_launches.csv"))
    shutil.copy(os.path.join(G, "r01_apply_flags.json"), os.path.join(P, "r01_apply_flags.json"))
for f in ("r01_bench_line.json", "r01_rows.json", "r01_red_peak.json", "r01_fused.json", "r01_direction_rotate.json",
          "r01_gcf.json"):
    shutil.copy(os.path.join(G, f), os.path.join(P, f))
shares = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_shares.py"), os.path.join(G, "r01_launches.csv")],
                        stdout=subprocess.PIPE, text=True).stdout
open(os.path.join(P, "r01_launch_shares.md"), "w").write(
    "# Kernel shares, `ncu --metrics gpu__time_duration.sum --clock-control none -c 400` of `python bench.py --steps 2 "
    "--warmup 3 --no-cpu-baseline`\n\n(5 device-resident steps followed by 5 host-buffer (e2e) steps, which launch the same "
    "kernels on 8 time chunks; per-launch times are cold-cache and serialised -- compare shares)\n\n" + shares)
get = raw_metrics("r01_window.ncu-rep")
rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
json.dump({"std_grid_dram_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
           "std_grid_red_sectors_per_launch": int(get("l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum")),
           "std_grid_red_instructions_per_launch": int(get("smsp__inst_executed_op_global_red.sum")),
           "source": "ncu --set full --clock-control none, profiles/r01_std_grid_window_f32_continuum.txt, one launch of "
                     "bench.py's gridding kernel (std_grid_window_kernel, C2, fp32, continuum)"},
          open(os.path.join(P, "r01_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, "r01_traffic.json")).read())
