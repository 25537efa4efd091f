import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
frac=float(sys.argv[2]) if len(sys.argv)>2 else 0.5
hdr=None; data=[]
for r in rows:
    if hdr is None:
        if r and r[0]=='ID': hdr=r
        continue
    data.append(r)
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.defaultdict(lambda:[0,0.0])
for r in data[int(len(data)*frac):]:
    n=r[ki][:70]; v=float(r[vi].replace(',',''))
    if r[ui]=='ns': v/=1e6
    elif r[ui]=='us': v/=1e3
    agg[n][0]+=1; agg[n][1]+=v
tot=sum(v[1] for v in agg.values())
print("total ms", round(tot,3), "launches", sum(v[0] for v in agg.values()))
for n,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[3]) if len(sys.argv)>3 else 14]:
    print(f"{t:9.3f} ms {c:5d}  {t/c*1e3:9.1f} us  {n}")
