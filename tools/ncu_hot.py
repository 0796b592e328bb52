"""Per-region breakdown of an ncu source page: executed instructions and stall samples between
barriers.  usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --launch-skip N --launch-count 1 > f.csv;
python tools/ncu_hot.py f.csv"""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address'][0]
hdr=rows[hi]
ia=hdr.index('Address'); isrc=hdr.index('Source'); iex=hdr.index('Instructions Executed'); ismp=hdr.index('# Samples')
data=[]
for r in rows[hi+1:]:
    if len(r)>iex:
        try: data.append((r[ia], r[isrc].strip(), int(r[iex]), int(r[ismp])))
        except ValueError: pass
tot=sum(d[2] for d in data); tots=sum(d[3] for d in data)
print('total warp-inst', tot, 'samples', tots)
seg=[];cur=[0,0,0,None]
for i,(a,s,e,sm) in enumerate(data):
    if cur[3] is None: cur[3]=i
    cur[0]+=e;cur[1]+=sm;cur[2]+=1
    if 'BAR.SYNC' in s or 'DEPBAR' in s:
        seg.append((cur[3],i,cur[0],cur[1],cur[2],s));cur=[0,0,0,None]
seg.append((cur[3],len(data)-1,cur[0],cur[1],cur[2],'END'))
for sg in seg: print(f"idx {sg[0]:5d}-{sg[1]:5d} ninstr {sg[4]:5d} exec {sg[2]/tot*100:6.2f}% samples {sg[3]/tots*100:6.2f}%  ends:{sg[5][:50]}")
if len(sys.argv)>2:
    a,b=int(sys.argv[2]),int(sys.argv[3])
    for i in range(a,b+1):
        d=data[i]; print(f"{i:5d} {d[2]:10d} {d[3]:6d}  {d[1][:90]}")
