"""Per-CUDA-source-line sample shares from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`.
Usage: python profiles/analyze_ncu_cuda_lines.py file.csv n_warps [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
W = float(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur = hdr = None
agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split('/')[-1]; continue
    if len(r) > 2 and r[0] == "Line No":
        hdr = r; iS = hdr.index('# Samples'); iE = hdr.index('Instructions Executed'); continue
    if hdr is None or len(r) < len(hdr) or r[0] == '':
        continue
    try:
        ln = int(r[0]); s = int(r[iS]); e = int(r[iE])
    except ValueError:
        continue
    agg[(cur, ln)] = (s, e, r[1].strip()[:100])
tot = sum(v[0] for v in agg.values())
print("total samples", tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0][:16]:16s}:{k[1]:4d} {v[0] / tot:6.2%} ex/warp={v[1] / W:7.0f}  {v[2]}")
