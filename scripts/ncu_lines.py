"""Aggregate an ncu source page (sass,cuda) per CUDA source line: python scripts/ncu_lines.py rep.ncu-rep [min_pct]"""
import csv, subprocess, sys, collections
rep=sys.argv[1]; thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.8
txt=subprocess.run(["ncu","-i",rep,"--page","source","--print-source","sass,cuda","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(txt.splitlines()))
agg=collections.OrderedDict(); cur_file=None; H=None
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur_file=r[1].split("/")[-1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": H=r; ii=H.index("Instructions Executed"); wi=H.index("Warp Stall Sampling (All Samples)"); continue
    if H is None or len(r)<=ii: continue
    if not r[0].strip(): continue   # SASS-only rows (no CUDA line): they duplicate the per-line totals
    key=(cur_file,r[0],r[1].strip())
    a=agg.setdefault(key,[0,0,collections.Counter()])
    try:
        a[0]+=int(r[ii]); a[1]+=int(r[wi])
    except ValueError: continue
    for j,h in enumerate(H):
        if h.startswith("stall_") and j<len(r) and r[j].isdigit() and int(r[j])>0: a[2][h[6:]]+=int(r[j])
ti=sum(a[0] for a in agg.values()); tw=sum(a[1] for a in agg.values())
print("total warp-instructions",ti,"stall samples",tw)
for (f,ln,src),a in agg.items():
    pi=a[0]/ti*100; pw=a[1]/tw*100
    if pi>=thr or pw>=thr:
        top=",".join(f"{k}:{v/tw*100:.1f}" for k,v in a[2].most_common(3))
        print(f"{f[:18]:18s}:{ln:>4s} inst={pi:5.1f}% stall={pw:5.1f}% [{top}]  {src[:90]}")
