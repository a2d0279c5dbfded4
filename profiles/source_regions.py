"""Sums the samples of an `ncu -i X.ncu-rep --page source --csv` page between the kernel's CTA barriers (one line per region:
SASS rows, samples, share, largest per-instruction execution count, top stall reasons).
Usage: python profiles/source_regions.py src.csv [window]   (window > 0: fixed windows of that many instructions instead)"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
win = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr, data = rows[1], rows[2:]
iS, iE, iSrc = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'not_issued' not in h.lower() and 'Not Issued' not in h]
tot = sum(int(r[iS]) for r in data)
print(f"# {rows[0][1][:90]}: {len(data)} SASS rows, {tot} samples")
if win:
    cuts = list(range(win - 1, len(data), win))
else:
    cuts = [i for i, r in enumerate(data) if 'BAR.SYNC' in r[iSrc]]
prev = 0
for b in cuts + [len(data) - 1]:
    w = data[prev:b + 1]
    if not w:
        continue
    sm = sum(int(r[iS]) for r in w)
    st = collections.Counter()
    for r in w:
        for i, h in stall_cols:
            st[h] += int(r[i] or 0)
    top = ", ".join(f"{h[6:]}={v / max(sm, 1):.0%}" for h, v in st.most_common(4))
    ops = collections.Counter()
    for r in w:
        t = r[iSrc].split()
        ops[(t[1] if t and t[0].startswith('@') and len(t) > 1 else (t[0] if t else '?')).split('.')[0]] += int(r[iS])
    hot = ", ".join(f"{k}={v / max(sm, 1):.0%}" for k, v in ops.most_common(3))
    print(f"rows {prev:5d}-{b:5d}  samples {sm:6d} {sm / tot:6.1%}  max exec {max(int(r[iE]) for r in w):8d} | stalls: {top} | sampled at: {hot}")
    prev = b + 1
