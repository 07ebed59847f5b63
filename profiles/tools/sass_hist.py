import csv, gzip, sys, re, collections
path=sys.argv[1]; pat=sys.argv[2] if len(sys.argv)>2 else ''
op=gzip.open if path.endswith('.gz') else open
kern=None; seen=collections.Counter()
hist={}; samples={}
with op(path,'rt') as f:
    r=csv.reader(f)
    for row in r:
        if not row: continue
        if row[0]=='Kernel Name':
            kern=re.sub(r'void pcd::pcd_kernel<pcd::K|, pcd::\w+>\(.*','',row[1]); seen[kern]+=1; kern=f"{kern}#{seen[kern]}"; hist[kern]=collections.Counter(); samples[kern]=collections.Counter(); continue
        if row[0]=='Address': hdr=row; continue
        src=row[1].strip()
        m=re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)',src)
        opc=m.group(2) if m else src
        base=opc.split('.')[0]
        if base in('LDS','STS','LDG','STG'):
            base=opc if ('128' in opc or '64' in opc) else base
            base='.'.join([p for p in base.split('.') if p in('LDS','STS','LDG','STG','128','64')])
        try: n=int(row[5]); s=int(row[4])
        except: continue
        hist[kern][base]+=n; samples[kern][base]+=s
for k in hist:
    if pat and pat not in k: continue
    tot=sum(hist[k].values()); ts=sum(samples[k].values())
    print(f"== {k}: warp-inst {tot/1e6:.2f}M samples {ts}")
    for o,n in hist[k].most_common(22):
        print(f"   {o:14s} {n/1e6:8.2f}M {100*n/tot:5.1f}%   samples {100*samples[k][o]/max(ts,1):5.1f}%")
