"""Per-kernel SASS opcode counts of the built library (cuobjdump -sass): the mnemonics that prove which kernels are on the
Blackwell paths -- UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG (TMA load), UTCBAR (tcgen05.commit),
HMMA (mma.sync), MUFU, LDL / STL (register spills), plus registers per thread.  MULTIMEM counts multimem.ld_reduce / .red
only: `multimem.st` is an ordinary system-scope STG in SASS (`STG.E.128.STRONG.SYS` in xrank_push_kernel; the switch
replicates it by address), which tests/test_sass_cpu.py checks.
    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections, hashlib, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "univst_b200", "libunivst_b200.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "HMMA", "MUFU", "FFMA2", "LDL", "STL", "SYNCS", "MULTIMEM", "RED", "ATOM"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
regs = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+)", res):
    regs[m.group(1)] = int(m.group(2))
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                counts[cur][o] += 1
print(f"library: univst_b200/libunivst_b200.so  sha256 {hashlib.sha256(open(LIB, 'rb').read()).hexdigest()[:16]}  (nvcc -gencode arch=compute_100a,code=sm_100a)")
print(f"{'kernel':100s} {'regs':>5s} {'instr':>7s} " + " ".join(f"{o:>8s}" for o in OPS))
tot = collections.Counter()
for fn, c in sorted(counts.items(), key=lambda kv: demangle(kv[0])):
    name = re.sub(r"\(.*", "", demangle(fn)).replace("void ", "").replace("uv::", "")
    print(f"{name[:100]:100s} {regs.get(fn, 0):5d} {c['_total']:7d} " + " ".join(f"{c[o]:8d}" for o in OPS))
    tot.update(c)
print(f"{'TOTAL':100s} {'':5s} {tot['_total']:7d} " + " ".join(f"{tot[o]:8d}" for o in OPS))
