import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
frames=float(sys.argv[2]); thr=float(sys.argv[3]) if len(sys.argv)>3 else 3
start=next(i for i,r in enumerate(rows) if r and r[0]=="Address")
hdr=rows[start]; data=[r for r in rows[start+1:] if len(r)>10]
iS=hdr.index('Source'); iE=hdr.index('Instructions Executed'); iSm=hdr.index('# Samples')
prev=None; st=0; seg=[]
for i,r in enumerate(data):
    n=int(r[iE] or 0)
    if prev is None or abs(n-prev)>0.02*max(n,prev,1):
        if prev is not None: seg.append((st,i))
        st=i
    prev=n
seg.append((st,len(data)))
print(len(data),'instr',len(seg),'regions')
for a,b in seg:
    tot=sum(int(r[iE] or 0) for r in data[a:b]); smp=sum(int(r[iSm] or 0) for r in data[a:b])
    if tot/frames<thr: continue
    ops=collections.Counter()
    for r in data[a:b]:
        s=r[iS].strip()
        if s.startswith('@'): s=s.split(None,1)[1]
        ops[s.split()[0].split('.')[0]]+=1
    n=int(data[a][iE])
    print(f'[{a:5d},{b:5d}) len {b-a:4d} exec/frame {n/frames:6.3f} -> {tot/frames:6.1f} inst/frame samples {smp:6d}', dict(ops.most_common(7)))
