#!/usr/bin/env python
"""Per-kernel SASS opcode counts of the in-tree library: tensor-pipe, copy-engine and barrier instructions.

    python scripts/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fastoverlap_b200", "lib")
OPS = ["DMMA", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "LDTM", "STTM", "UTCATOMSWS", "BAR", "REDUX", "RED", "ATOMS", "DFMA", "DADD", "DMUL"]
out = collections.OrderedDict()
for obj in sorted(f for f in os.listdir(LIB) if f.endswith(".o")):
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(LIB, obj)], capture_output=True, text=True).stdout
    name = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0].replace("void ", "")
            out[(obj, name)] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    out[(obj, name)][o] += 1
            out[(obj, name)]["total"] += 1
print("# SASS summary of fastoverlap_b200/lib/*.o (sm_100a): static instruction counts per kernel")
print("# DMMA = mma.sync.m8n8k4.f64 (FP64 tensor pipe); UBLKCP = cp.async.bulk (copy engine); LDGSTS = cp.async;")
print("# SYNCS = mbarrier; LDTM / STTM = tcgen05.ld / st (tensor memory); tcgen05.mma has no FP64 kind, so no UTCMMA here")
print("%-18s %-58s %7s " % ("object", "kernel", "instrs") + " ".join("%7s" % o for o in OPS))
for (obj, name), c in out.items():
    if c["total"] < 50:
        continue
    print("%-18s %-58s %7d " % (obj, name[:58], c["total"]) + " ".join("%7d" % c[o] for o in OPS))
