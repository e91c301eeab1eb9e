"""Summarise an ncu report's source page: per-SASS-instruction stall samples of the first kernel.
usage: ncu_src.py report.ncu-rep [min_samples] [start end]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
minS = int(sys.argv[2]) if len(sys.argv) > 2 else 600
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
print(rows[0][1][:100])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
out = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    out.append(r)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in out)
print("instructions", len(out), "total samples", tot)
a, b = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (0, len(out))
for i in range(a, b):
    r = out[i]
    s = int(r[ix["# Samples"]])
    if s < minS:
        continue
    top = sorted(((st[6:], int(r[ix[st]])) for st in stalls), key=lambda x: -x[1])[:3]
    print(i, s, f"{100 * s / tot:.2f}%", r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:64], top)
