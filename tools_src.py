import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot=0; per=[]
agg={s:0 for s in stalls}
for r in rows[2:]:
    if r and r[0] in ("Address","Kernel Name"): continue
    if len(r)<len(hdr): continue
    n=int(r[idx['# Samples']] or 0); tot+=n
    for s in stalls: agg[s]+=int(r[idx[s]] or 0)
    per.append((n, r[idx['Source']][:110], {s:int(r[idx[s]] or 0) for s in stalls if int(r[idx[s]] or 0)>0}, r[idx['Instructions Executed']]))
print("total samples", tot)
print({k:v for k,v in sorted(agg.items(), key=lambda kv:-kv[1]) if v>0})
top=int(sys.argv[2]) if len(sys.argv)>2 else 25
for n,src,st,ie in sorted(per,key=lambda x:-x[0])[:top]:
    print("%6d %5.1f%%  %-110s  exec=%s %s" % (n, 100.0*n/tot, src, ie, dict(sorted(st.items(), key=lambda kv:-kv[1])[:3])))
