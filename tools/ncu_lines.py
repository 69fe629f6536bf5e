#!/usr/bin/env python3
"""Per-source-line view of an ncu capture: joins `ncu --page source --csv` (one row per SASS instruction) with
the line table of the cubin (`nvdisasm -g`), and prints the lines that carry most samples / instructions /
shared-memory wavefronts.  usage: tools/ncu_lines.py <report.ncu-rep> <kernel-regex> <launch-index> [lib.so] [top]"""
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kre, idx = sys.argv[1], sys.argv[2], sys.argv[3]      # idx: launch index among the matching kernels, or "max" = the longest one
lib = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(__file__), "..", "figbird_b200", "_build", "libfigbird_b200.so")
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL, check=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin") and f.count("-") == 0 or f.startswith("fb_engine.")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
# address -> line for the kernel whose section name matches
line_of, cur_fn, cur_line = {}, None, 0
for l in dis.split("\n"):
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        cur_fn = m.group(1); continue
    m = re.search(r'//## File ".*?", line (\d+)', l)
    if m:
        cur_line = int(m.group(1)); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m and cur_fn and re.search(kre, cur_fn):
        line_of.setdefault(cur_fn, {})[int(m.group(1), 16)] = cur_line
if idx == "max":
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    rr = list(csv.reader(raw.split("\n")))
    h = rr[0]; dcol = h.index("gpu__time_duration.sum"); ncol = h.index("Kernel Name")
    cand = [(float(r[dcol]), k + 1) for k, r in enumerate([r for r in rr[2:] if len(r) == len(h) and re.search(kre, r[ncol])])]
    idx = str(max(cand)[1])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", "::regex:%s:%s" % (kre, idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
rows = list(csv.reader(out.split("\n")))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
kname = rows[0][1] if rows and len(rows[0]) > 1 else ""
templ = "ILb1" if "(bool)1" in kname or "<1>" in kname else ("ILb0" if "(bool)0" in kname or "<0>" in kname else "")
fn = next((f for f in line_of if templ in f), next(iter(line_of)))
lo = line_of[fn]
iA, iI, iS, iW, iT = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared"), hdr.index("Thread Instructions Executed")
base = None
agg = {}
for r in rows[hi + 1:]:
    if len(r) <= iW or not r[iA]:
        continue
    try:
        a = int(r[iA], 16)
    except ValueError:
        continue
    if base is None:
        base = a
    ln = lo.get(a - base, -1)
    d = agg.setdefault(ln, [0, 0, 0, 0])
    d[0] += int(r[iI] or 0); d[1] += int(r[iS] or 0); d[2] += int(r[iW] or 0); d[3] += int(r[iT] or 0)
src = open(os.path.join(os.path.dirname(__file__), "..", "figbird_b200", "csrc", "fb_engine.cu")).read().split("\n")
ti, ts, tw = (sum(d[k] for d in agg.values()) or 1 for k in range(3))
print("kernel %s  (%s)\ntotal warp instructions %d, samples %d, shared wavefronts %d" % (kname, fn[-40:], ti, ts, tw))
print("%6s %7s %7s %7s %6s  %s" % ("line", "inst%", "samp%", "smem%", "thr/in", "source"))
for ln, d in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%6d %7.2f %7.2f %7.2f %6.1f  %s" % (ln, 100.0 * d[0] / ti, 100.0 * d[1] / ts, 100.0 * d[2] / tw, d[3] / max(d[0], 1), src[ln - 1].strip()[:120] if 0 < ln <= len(src) else "?"))
