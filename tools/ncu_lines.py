#!/usr/bin/env python
"""Stall samples and executed instructions of one kernel of an .ncu-rep, aggregated by CUDA source line (read here, no GPU):

    python tools/ncu_lines.py gpurun_out/x.ncu-rep <kernel regex> <object.o> <mangled-name substring> [top]

The .ncu-rep's SASS page carries no line numbers in --csv mode; they come from `nvdisasm -g` of the same object file
(instruction order is identical), so the object must be the build that was profiled."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, pat, obj, sub = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and sub in l][0]
cur, seq = None, []
for l in dis[start + 1:]:
    if l.startswith("//-----"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+\S", l):
        seq.append(cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", "::regex:%s:1" % pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)][:len(seq)]
assert len(data) == len(seq), (len(data), len(seq))
by = collections.defaultdict(lambda: [0.0, 0.0])
te = ts = 0.0
for r, c in zip(data, seq):
    e, s = float(r[ix["Instructions Executed"]] or 0), float(r[ix["# Samples"]] or 0)
    by[c][0] += e
    by[c][1] += s
    te += e
    ts += s
print("warp instructions %d, samples %d" % (te, ts))
cache = {}
for k, (e, s) in sorted(by.items(), key=lambda x: -x[1][1])[:top]:
    txt = ""
    if k and os.path.exists(k[0]):
        cache.setdefault(k[0], open(k[0]).read().split("\n"))
        txt = cache[k[0]][k[1] - 1].strip()[:100]
    print("%s:%s  %5.1f%% instr %5.1f%% samples  %s" % (os.path.basename(k[0]) if k else None, k[1] if k else "", 100 * e / te, 100 * s / ts, txt))
