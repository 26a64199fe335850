"""Summarise an ncu report per CUDA source line: share of executed instructions and of stall samples.
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kern = ["--kernel-name", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
out = subprocess.run(["ncu", "-i", rep] + kern + ["--page", "source", "--print-source", "cuda,sass", "--csv"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, fname, lines = None, "", []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and len(r) == len(hdr) and (not lines or len(r) == len(lines[0][1])):
        lines.append((fname, r))
i_inst = hdr.index("Instructions Executed")
i_thr = hdr.index("Thread Instructions Executed")
i_samp = hdr.index("# Samples")
i_src = 1
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot_i = sum(float(r[i_inst] or 0) for _, r in lines)
tot_s = sum(float(r[i_samp] or 0) for _, r in lines)
tot_t = sum(float(r[i_thr] or 0) for _, r in lines)
print("total warp instructions %.4g   thread instr %.4g (avg active lanes %.1f)   samples %d" % (tot_i, tot_t, tot_t / tot_i, tot_s))
agg = {}
for i, h in stall_cols:
    agg[h] = sum(float(r[i] or 0) for _, r in lines if i < len(r))
print("stall mix: " + ", ".join("%s %.1f%%" % (h[6:], 100 * v / tot_s) for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
lines.sort(key=lambda fr: -float(fr[1][i_inst] or 0))
print("%6s %6s %5s  %s" % ("inst%", "samp%", "lanes", "line"))
for f, r in lines[:top]:
    inst = float(r[i_inst] or 0)
    lanes = float(r[i_thr] or 0) / inst if inst else 0
    print("%5.1f%% %5.1f%% %5.1f  %s:%s  %s" % (100 * inst / tot_i, 100 * float(r[i_samp] or 0) / tot_s, lanes, f, r[0], r[i_src].strip()[:100]))
