import csv,sys,collections
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); ib=hdr.index('Block Size')
d=collections.OrderedDict()
for r in rows[1:]:
    k=r[ik][:44]+r[ib]; d.setdefault(k,[]).append(float(r[iv].replace(',',''))/1e6)
tot=0
for k,v in d.items():
    v=sorted(v); m=v[len(v)//2]; tot+=m; print(f"{k:60s} n={len(v):3d} median {m:.3f} ms  min {v[0]:.3f}")
print("sum of medians", round(tot,3))
