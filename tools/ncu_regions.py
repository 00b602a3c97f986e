"""Groups the per-line rows of tools/ncu_lines.py by enclosing function (and by the phase markers
inside the MC kernel).

usage: ncu_regions.py <ncu_lines output file>
"""
import collections
import os
import re
import sys

root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "crystal-packing_b200", "csrc")
MARKERS = [("Draw until a proposal", "mc: draw loop"), ("// phase A", "mc: phase A"), ("// phase B", "mc: phase B"),
           ("// phase C", "mc: phase C"), ("one entry per lane: the certificate", "overlap_warp: certificate"),
           ("component pairs without a certificate", "overlap_warp: queue push"),
           ("const unsigned wanting", "overlap_warp: pool round")]
region_of = {}
for f in os.listdir(root):
    if not f.endswith((".cuh", ".h")):
        continue
    cur = f
    table = {}
    for i, ln in enumerate(open(os.path.join(root, f)).read().splitlines(), 1):
        m = re.match(r"^(?:static\s+)?(?:__device__|__global__|__host__|CPF_HD|CP_HD|CPF_CX|template|inline|constexpr|struct)\b.*?(\w+)\s*\(", ln)
        if m and not ln.startswith("template"):
            cur = f.split(".")[0] + ":" + m.group(1)
        m2 = re.match(r"^(mc_thread_kernel|score_thread_kernel|replay_thread_kernel)\(", ln)
        if m2:
            cur = f.split(".")[0] + ":" + m2.group(1)
        for key, name in MARKERS:
            if key in ln:
                cur = name
        table[i] = cur
    region_of[f] = table
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for ln in open(sys.argv[1]):
    m = re.match(r"(\S+)\s+(\d+)\s+([\d.]+)%\s+([\d.]+)/move\s+lanes\s+([\d.]+)\s+samples\s+([\d.]+)%", ln)
    if not m:
        continue
    f, l, pm, lanes, smp = m.group(1), int(m.group(2)), float(m.group(4)), float(m.group(5)), float(m.group(6))
    name = region_of.get(f, {}).get(l, "other:" + f)
    a = agg[name]
    a[0] += pm
    a[1] += pm * lanes
    a[2] += smp
tot = sum(v[0] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:44s} {v[0]:7.2f}/move {v[0] / tot * 100:5.1f}%  lanes {v[1] / max(v[0], 1e-9):5.1f}  samples {v[2]:5.1f}%")
print(f"total {tot:.2f} warp-instructions per move")
