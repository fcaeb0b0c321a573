"""Opcode histogram per kernel of the built library (cuobjdump -sass): the evidence that the hot kernels are
Blackwell-native (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP / UTMALDG / UTMASTG = TMA, UTCBAR =
tcgen05.commit, SYNCS = mbarrier).  Usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "ultra_pytorch_b200", "lib", "libultra_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "UTMALDG", "UTMASTG", "UTMACMDFLUSH", "SYNCS",
        "USETMAXREG", "HMMA", "FFMA", "MUFU", "LDG", "STG", "LDS", "STS", "ATOM", "RED", "BAR", "LDL", "STL"]
kern, counts, order = None, {}, []
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        counts[kern] = collections.Counter()
        order.append(kern)
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and kern:
        op = m.group(1)
        counts[kern]["_total"] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[kern][k] += 1
                break
print("library: %s" % os.path.relpath(lib, ROOT))
print("%-58s %7s  %s" % ("kernel", "instrs", "selected opcodes"))
for k in order:
    c = counts[k]
    sel = "  ".join("%s=%d" % (q, c[q]) for q in KEYS if c[q])
    print("%-58s %7d  %s" % (k[:58], c["_total"], sel))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("\nTOTAL  " + "  ".join("%s=%d" % (q, tot[q]) for q in KEYS if tot[q]))
