"""Summarise an ncu gpu__time_duration launch list: python scripts/launch_summary.py gpurun_out/launches.csv"""
import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]; ki=H.index('Kernel Name'); vi=H.index('Metric Value'); ui=H.index('Metric Unit')
agg=collections.defaultdict(list)
for r in rows[hdr+1:]:
    if len(r)<=vi: continue
    v=float(r[vi].replace(',','')); u=r[ui]
    v = v/1e3 if u in('nsecond','ns') else v*1e3 if u in('msecond','ms') else v
    agg[r[ki].replace('unnamed>::','').replace('void ','')[:60]].append(v)
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    print(f"{k:62s} n={len(v):3d} avg={sum(v)/len(v):9.1f}us share={sum(v)/tot*100:5.1f}%")
