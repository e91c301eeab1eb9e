"""Aggregate an ncu report's source page: stall samples by reason and by opcode (first kernel in the report).
usage: ncu_stalls.py report.ncu-rep"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
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
by_reason = collections.Counter()
by_op = collections.Counter()
by_op_reason = collections.defaultdict(collections.Counter)
exec_by_op = collections.Counter()
for r in out:
    src = r[ix["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = op.split(".")[0]
    exec_by_op[op] += int(r[ix["Instructions Executed"]] or 0)
    for st in stalls:
        v = int(r[ix[st]] or 0)
        by_reason[st[6:]] += v
        by_op[op] += v
        by_op_reason[op][st[6:]] += v
print("total samples", tot)
print("-- by stall reason")
for k, v in by_reason.most_common(14):
    print(f"  {k:22s} {v:8d} {100.0 * v / tot:5.1f}%")
print("-- by opcode (samples at that instruction = waiting to issue it)")
for k, v in by_op.most_common(22):
    top = ", ".join(f"{a}:{b}" for a, b in by_op_reason[k].most_common(3))
    print(f"  {k:10s} {v:8d} {100.0 * v / tot:5.1f}%  warp-instr executed {exec_by_op[k]:10d}   {top}")
