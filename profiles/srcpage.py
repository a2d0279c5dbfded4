"""Aggregates an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line:
python profiles/srcpage.py export.csv source.cu warp_steps [min_instr]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
srcname = sys.argv[2].split('/')[-1]
src = open(sys.argv[2]).read().splitlines()
W = float(sys.argv[3]); thr = float(sys.argv[4]) if len(sys.argv) > 4 else 4.0
def I(x):
    try: return int(x)
    except Exception: return 0
cur = '?'; agg = collections.OrderedDict(); hdr = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or not r[0].isdigit(): continue
    d = dict(zip(hdr, r))
    key = (cur.split('/')[-1], int(r[0]))
    a = agg.setdefault(key, [0, 0, collections.Counter()])
    a[0] += I(d.get('# Samples')); a[1] += I(d.get('Instructions Executed'))
    for k, v in d.items():
        if k.startswith('stall_') and 'Not Issued' not in k and I(v): a[2][k] += I(v)
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
print(f"samples {ts}  instructions {ti}  per warp-step {ti / W:.1f}")
for k, v in sorted(agg.items(), key=lambda kv: (kv[0][0] != srcname, kv[0][1])):
    if v[1] / W > thr or v[0] > 0.012 * ts:
        line = src[k[1] - 1].strip()[:64] if k[0] == srcname else k[0]
        print(f"{k[1]:4d} {v[1] / W:6.1f} i/ws {100 * v[0] / ts:5.1f}% {dict(v[2].most_common(2))} | {line}")
