import csv, collections, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ik][:50]].append(float(r[iv].replace(',','')))
    except: pass
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])): print(f'{k:52s} n={len(v):3d} mean={sum(v)/len(v)/1000:9.1f} us max={max(v)/1000:8.1f} total={sum(v)/1e6:7.2f} ms share={100*sum(v)/tot:5.1f}%')
print('total ms', tot/1e6)
