import csv, sys, collections, re
def opmix(path, top=25):
    rows=list(csv.reader(open(path)))
    hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}
    ops=collections.Counter(); tot=0; stall=collections.Counter()
    for r in rows[2:]:
        if len(r)<len(hdr) or r[0]=="Address": continue
        src=r[idx["Source"]].strip()
        m=re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
        if not m: continue
        op=m.group(2).split('.')[0]
        n=int(r[idx["Instructions Executed"]] or 0)
        ops[op]+=n; tot+=n
        stall[op]+=int(r[idx["Warp Stall Sampling (All Samples)"]] or 0)
    print("total warp-inst", tot)
    st=sum(stall.values())
    for op,n in ops.most_common(top):
        print(f"  {op:10s} {n:12d} {100*n/tot:5.1f}%   stall-samples {100*stall[op]/max(st,1):5.1f}%")
opmix(sys.argv[1])
