import csv, sys
lines=[l for l in open(sys.argv[1]) if l.startswith('"')]
rd=csv.reader(lines); hdr=next(rd); ix={h:i for i,h in enumerate(hdr)}
rows=[(int(r[ix['ID']]), r[ix['Kernel Name']], float(r[ix['Metric Value']]), r[ix['Grid Size']]) for r in rd]
# last step: between the last two multi_tensor_apply (SGD) clusters -> use the last N rows after the previous "FillFunctor<unsigned char>" (flush)
fl=[i for i,r in enumerate(rows) if 'FillFunctor<unsigned char>' in r[1]]
a,b=fl[-2]+1, fl[-1]
step=rows[a:b]
tot=sum(r[2] for r in step)
agg={}
for _,n,t,g in step:
    k=n.split('(')[0][:90]
    c=agg.setdefault(k,[0.0,0]); c[0]+=t; c[1]+=1
print('%d launches, %.1f us'%(len(step),tot/1e3))
for k,(t,n) in sorted(agg.items(), key=lambda kv:-kv[1][0])[:40]:
    print('%8.1f us %5.1f%% x%3d %s'%(t/1e3,100*t/tot,n,k))
