"""Summarise `ncu -i X.ncu-rep --page source --csv` output: hot SASS regions, opcode mix, stall reasons.
Usage: python profiles/analyze_ncu_source.py src.csv n_warps [min_exec_per_warp]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
W = float(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 100
hdr, data = rows[1], rows[2:]
iS, iE, iSrc = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot_s = sum(int(r[iS]) for r in data)
print(f"kernel: {rows[0][1][:80]}  total samples {tot_s}  instr/warp {sum(int(r[iE]) for r in data) / W:.0f}")
tot_st = collections.Counter()
for r in data:
    for i, h in stall_cols:
        tot_st[h] += int(r[i] or 0)
print("stalls:", ", ".join(f"{h[6:]}={v / tot_s:.1%}" for h, v in tot_st.most_common(8)))
segs, seg = [], None
for i, r in enumerate(data):
    if int(r[iE]) / W >= thr:
        if seg is None:
            seg = [i, i]
        seg[1] = i
    elif seg is not None and i - seg[1] > 12:
        segs.append(seg); seg = None
if seg:
    segs.append(seg)
for a, b in segs:
    s = sum(int(r[iS]) for r in data[a:b + 1]); e = sum(int(r[iE]) for r in data[a:b + 1]) / W
    if s / tot_s < 0.01:
        continue
    ops, st = collections.Counter(), collections.Counter()
    for r in data[a:b + 1]:
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[iSrc])
        ops[m.group(2) if m else '?'] += int(r[iE]) / W
        for i, h in stall_cols:
            st[h] += int(r[i] or 0)
    print(f"\nSASS lines {a}-{b}: {b - a + 1} instr, exec/warp {e:.0f}, samples {s / tot_s:.1%}")
    print("   ops  :", ", ".join(f"{k}={v:.0f}" for k, v in ops.most_common(10)))
    print("   stall:", ", ".join(f"{h[6:]}={v / max(s, 1):.0%}" for h, v in st.most_common(6)))
