#!/usr/bin/env python
"""Instruction counts of a kernel per barrier-separated segment, from `ncu -i rep --page source --csv` (SASS view).

    python tools/sass_segments.py gpurun_out/fused_source.csv [kernel-name-substring]
"""
import collections
import csv
import sys

path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
kern, hdr, data = None, None, {}
for row in csv.reader(open(path)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        kern, hdr = row[1], None
        data[kern] = []
    elif row[0] == "Address":
        hdr = row
    elif kern and hdr:
        data[kern].append(row)
for k, rows in data.items():
    if want not in k:
        continue
    print(k[:110], len(rows), "SASS instructions")
    iS, iI, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    iW, iWi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
    iBar = hdr.index("stall_barrier")
    seg, tot, tot_s = 0, 0, 0
    segs = collections.defaultdict(lambda: [0, 0, 0, 0, 0, collections.Counter(), 0])
    for row in rows:
        ins, sm = int(row[iI] or 0), int(row[iSm] or 0)
        words = row[iS].split()
        op = words[1] if words and words[0].startswith("@") and len(words) > 1 else (words[0] if words else "")
        s = segs[seg]
        s[0] += ins
        s[1] += sm
        s[2] += 1
        s[3] += int(row[iW] or 0)
        s[4] += int(row[iWi] or 0)
        s[5][op.split(".")[0]] += ins
        s[6] += int(row[iBar] or 0)
        tot += ins
        tot_s += sm
        if op.startswith("BAR"):
            seg += 1
    for sg, (ins, sm, n, w, wi, ops, bar) in sorted(segs.items()):
        print(f"seg {sg:2d}: sass {n:5d} inst {ins / 1e6:8.2f}M ({100 * ins / max(1, tot):4.1f}%) samples {sm:6d} ({100 * sm / max(1, tot_s):4.1f}%, barrier {bar:5d}) "
              f"smem wavefronts {w / 1e6:6.2f}M ideal {wi / 1e6:6.2f}M  top:", ", ".join(f"{o}:{c / 1e6:.1f}" for o, c in ops.most_common(7)))
    print(f"total {tot / 1e6:.1f}M warp instructions, {tot_s} samples")
