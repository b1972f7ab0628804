import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}
stall_cols=[h for h in hdr if h.startswith("stall_") or "Stall" in h]
recs=[]
for r in rows[2:]:
    if len(r)<len(hdr) or r[0]=="Address": continue
    n=int(r[idx["Warp Stall Sampling (All Samples)"]] or 0)
    recs.append((n,r))
tot=sum(n for n,_ in recs)
print("total samples", tot)
# aggregate reasons
agg={}
for n,r in recs:
    for h in hdr:
        if h.startswith("stall_") and not h.endswith("_not_issued"):
            try: agg[h]=agg.get(h,0)+int(r[idx[h]] or 0)
            except: pass
for h,v in sorted(agg.items(), key=lambda kv:-kv[1])[:12]:
    print(f"  {h:35s} {100*v/max(tot,1):5.1f}%")
recs.sort(key=lambda t:-t[0])
for n,r in recs[:int(sys.argv[2]) if len(sys.argv)>2 else 20]:
    reasons=[(h,int(r[idx[h]] or 0)) for h in hdr if h.startswith("stall_") and not h.endswith("_not_issued")]
    reasons=[x for x in sorted(reasons,key=lambda t:-t[1]) if x[1]>0][:3]
    print(f"{100*n/tot:5.1f}%  {r[idx['Source']].strip()[:70]:70s} {reasons}")
