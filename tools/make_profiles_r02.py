"""Turns the artefacts of tools/collect_profiles_r02.sh into the summaries tracked under profiles/r02_*."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
# Runs ON THE GPU BOX at the end of tools/collect_profiles_r02.sh (the .ncu-rep files together exceed what gpurun copies
# back): reads gpurun_out/*.ncu-rep, writes the text / json summaries into gpurun_out/r02_profiles/, which are then
# copied into profiles/ in the container (`cp gpurun_out/r02_profiles/* profiles/`).
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(G, "r02_profiles")
os.makedirs(P, exist_ok=True)


def summary(rep, out):
    if not os.path.exists(os.path.join(G, rep)):
        print("missing", rep)
        return
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(G, rep)],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    open(os.path.join(P, out), "w").write(txt)


def raw_metrics(rep, row=2):
    raw = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[row]
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9}

    def get(name):
        i = hdr.index(name)
        return float(r[i].replace(",", "")) * mult.get(units[i], 1)
    return get


summary("r02_window.ncu-rep", "r02_std_grid_window_f32_continuum.txt")
summary("r02_window_iw.ncu-rep", "r02_std_grid_window_fused_weights_f32.txt")
summary("r02_bluestein.ncu-rep", "r02_bluestein_9830.txt")
summary("r02_window_cube.ncu-rep", "r02_std_grid_window_f32_cube_9830.txt")
summary("r02_iw.ncu-rep", "r02_imaging_weight_kernels.txt")
if os.path.exists(os.path.join(G, "r02_iw.ncu-rep")):   # per-line tables of every kernel generation in that capture
    with open(os.path.join(P, "r02_imaging_weight_kernels.txt"), "a") as f:
        f.write("\n# iw_grid_kernel / iw_degrid_mlp_kernel: the general kernels (round 1); iw_grid_fast_kernel / iw_degrid_fast_kernel: "
                "the product path (round 2)\n")
        for k in ("iw_grid_kernel", "iw_grid_fast", "iw_degrid_mlp", "iw_degrid_fast"):
            f.write("\n### %s\n" % k)
            f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), os.path.join(G, "r02_iw.ncu-rep"),
                                    "16", k], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout)
summary("r02_aperture.ncu-rep", "r02_aperture_track_f32.txt")
summary("r02_aperture_bulk.ncu-rep", "r02_aperture_track_bulk_ring_f32.txt")
for f in ("r02_launches.csv", "r02_fft.json", "r02_fused_weights.json", "r02_rows.json"):
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(P, f))
shares = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_shares.py"), os.path.join(G, "r02_launches.csv")],
                        stdout=subprocess.PIPE, text=True).stdout
open(os.path.join(P, "r02_launch_shares.md"), "w").write(
    "# Kernel shares, `ncu --metrics gpu__time_duration.sum --clock-control none -c 200` of `python bench.py --steps 2 "
    "--warmup 3 --no-cpu-baseline --no-e2e --no-cube --no-extras --no-parity`\n\n(5 device-resident steps; per-launch times are "
    "cold-cache and serialised -- compare shares, not absolutes)\n\n" + shares)
get = raw_metrics("r02_window.ncu-rep")
rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
out = {"std_grid_kernel": "std_grid_window_kernel<float,complex,S=7,PP=2>",
       "std_grid_dram_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
       "std_grid_red_sectors_per_launch": int(get("l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum")),
       "std_grid_red_instructions_per_launch": int(get("smsp__inst_executed_op_global_red.sum")),
       "std_grid_issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
       "std_grid_note": "not HBM bound: %.0f %% of the issue slots are active on 16 warps per SM and the shared-memory pipe co-limits "
                        "it; FP32 floor 0.45 ms per launch (256 FMAs per sample, 128 FMA lanes per clock and SM); see DESIGN.md "
                        "section 4.1" % get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
       "source": "ncu --set full --clock-control none, profiles/r02_std_grid_window_f32_continuum.txt, one launch of bench.py's "
                 "gridding kernel (C2, fp32, continuum)"}
if os.path.exists(os.path.join(G, "r02_window_cube.ncu-rep")):
    gc = raw_metrics("r02_window_cube.ncu-rep")
    log = open(os.path.join(G, "r02_ncu_cube.log")).read()
    n = int([l for l in log.splitlines() if l.startswith("samples_per_launch")][0].split()[1])
    out["cube_red_sectors_per_sample"] = gc("l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum") / n
    out["cube_probe"] = {"samples_per_launch": n, "kernel_ms": gc("gpu__time_duration.sum") * 1e3,
                         "dram_bytes": int(gc("dram__bytes_read.sum") + gc("dram__bytes_write.sum")),
                         "what": "tools/probe_cube_chunk.py: 8 channels x 2 pol of the config-5 cube, 9830^2 planes, one launch"}
json.dump(out, open(os.path.join(P, "r02_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, "r02_traffic.json")).read())
