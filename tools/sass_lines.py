"""Static SASS instruction count per source line of one kernel (nvdisasm line table).

usage: sass_lines.py <object-or-cubin> <mangled kernel name> [top]
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

obj, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
if not obj.endswith(".cubin"):
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    obj = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", obj], capture_output=True, text=True).stdout.splitlines()
agg, ops, cur, inside, total = collections.Counter(), collections.defaultdict(collections.Counter), None, False, 0
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
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?(\S+)", ln)
    if m:
        agg[cur] += 1
        ops[cur][m.group(2).split(".")[0].rstrip(";")] += 1
        total += 1
print("total", total)
for key, n in agg.most_common(top):
    f, l = key if key else ("?", 0)
    print(f"{f:18s}{l:5d} {n:5d}  " + " ".join(f"{k}:{v}" for k, v in ops[key].most_common(5)))
