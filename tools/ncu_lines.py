"""Per-source-line view of an ncu report: joins the SASS rows of `ncu --page source --csv` with
nvdisasm's line table of the same kernel.

usage: ncu_lines.py <report.ncu-rep> <object-or-cubin> <mangled kernel name> <moves in launch> [top]
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, obj, kernel, moves = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
if not obj.endswith(".cubin"):
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    obj = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", obj], capture_output=True, text=True).stdout.splitlines()
lines, cur, inside = {}, None, False
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        inside = ("text." + kernel + " ") in ln.replace("-", " ")
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S+)", ln)
    if m:
        lines[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
base = int(data[0][ix["Address"]], 16)
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in data:
    off = int(r[ix["Address"]], 16) - base
    key = lines.get(off, ("?", 0))
    a = agg[key]
    a[0] += int(r[ix["Instructions Executed"]])
    a[1] += int(r[ix["Thread Instructions Executed"]])
    a[2] += int(r[ix["# Samples"]])
tot = sum(a[0] for a in agg.values())
ts = sum(a[2] for a in agg.values())
print(f"total warp-inst {tot}  per move {tot / moves:.1f}; rows by source line (inlined callee lines)")
srcs = {}
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    for d in ("crystal-packing_b200/csrc",):
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
        if os.path.exists(path):
            srcs.setdefault(path, open(path).read().splitlines())
            if 0 < l <= len(srcs[path]):
                text = srcs[path][l - 1].strip()[:80]
    print(f"{f:16s}{l:5d} {a[0] / tot * 100:5.1f}%  {a[0] / moves:7.2f}/move  lanes {a[1] / max(a[0], 1):5.1f}  "
          f"samples {a[2] / max(ts, 1) * 100:5.1f}%  {text}")
