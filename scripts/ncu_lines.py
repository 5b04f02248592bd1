#!/usr/bin/env python
"""Per-source-line stall samples of one kernel from an ncu report.

    python scripts/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [--top 40] [--lib path/to/lib.so]

`ncu --page source --csv` lists SASS with per-instruction samples but no line numbers; `nvdisasm -g`
of the same cubin lists the SASS with line markers.  Both are in address order, so the two listings
are joined by instruction index (opcode checked).  The .so must be the build the report was taken on.
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ncu_sass(report, kernel, launch=None):
    cmd = ["ncu", "-i", report, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel]
    if launch is not None:
        cmd += ["--launch-skip", str(launch), "--launch-count", "1"]
    txt = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    out, hdr, name = [], None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            if out:
                break  # first matching launch only
            name = r[1]
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            out.append(dict(zip(hdr, r)))
    return name, out


def disasm_lines(lib, mangled_hint):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
    res = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True,
                             text=True).stdout
        cur, line, fn = None, None, None
        for ln in txt.splitlines():
            m = re.match(r"\.text\.(\S+):", ln)
            if m:
                cur = m.group(1)
                res[cur] = []
                line = None
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                fn, line = os.path.basename(m.group(1)), int(m.group(2))
                continue
            if cur is None:
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", ln)
            if m:
                res[cur].append((fn, line, m.group(1).strip()))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--lib", default=os.path.join(ROOT, "fastoverlap_b200", "lib", "libfastoverlap_b200.so"))
    ap.add_argument("--src-context", action="store_true")
    args = ap.parse_args()
    name, sass = ncu_sass(args.report, args.kernel)
    if not sass:
        sys.exit("no kernel matching %r in %s" % (args.kernel, args.report))
    print("kernel:", name, " instructions:", len(sass))
    # demangled name -> find the text section with the same instruction count
    fns = disasm_lines(args.lib, name)
    short = re.search(r"::(\w+)", name).group(1) if "::" in name else name.split("(")[0]
    cands = [(k, v) for k, v in fns.items() if short in k and len(v) == len(sass)]
    if not cands:
        print("no text section with %d instructions for %s; candidates:" % (len(sass), short))
        for k, v in fns.items():
            if short in k:
                print("  ", len(v), k)
        sys.exit(1)
    # opcode check
    best = None
    for k, v in cands:
        ok = sum(1 for a, b in zip(sass, v) if a["Source"].split()[0].lstrip("@!UP0123456789 ") [:3]
                 == b[2].split()[0].lstrip("@!UP0123456789 ")[:3])
        if best is None or ok > best[0]:
            best = (ok, k, v)
    print("matched section:", best[1][:100], " opcode agreement %d/%d" % (best[0], len(sass)))
    per = collections.defaultdict(lambda: collections.Counter())
    tot = 0
    stall_cols = [c for c in sass[0] if c.startswith("stall_") and "Not Issued" not in c]
    for row, (fn, line, ins) in zip(sass, best[2]):
        s = int(row["# Samples"] or 0)
        tot += s
        key = (fn, line)
        per[key]["samples"] += s
        per[key]["inst"] += int(row["Instructions Executed"] or 0)
        for c in stall_cols:
            v = int(row[c] or 0)
            if v:
                per[key][c] += v
    src = {}
    print("total samples", tot)
    for (fn, line), c in sorted(per.items(), key=lambda kv: -kv[1]["samples"])[:args.top]:
        stalls = ", ".join("%s %d" % (k[6:], v) for k, v in c.most_common() if k.startswith("stall_"))[:90]
        text = ""
        if fn:
            if fn not in src:
                for d in ("fastoverlap_b200/csrc", "scripts"):
                    p = os.path.join(ROOT, d, fn)
                    if os.path.exists(p):
                        src[fn] = open(p).read().splitlines()
                src.setdefault(fn, [])
            if line and line <= len(src[fn]):
                text = src[fn][line - 1].strip()[:70]
        print("%5.1f%% %7d inst %9d  %s:%s  %-70s | %s" % (100.0 * c["samples"] / max(tot, 1), c["samples"],
                                                         c["inst"], fn, line, text, stalls))


if __name__ == "__main__":
    main()
