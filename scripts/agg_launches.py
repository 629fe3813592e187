import csv, collections, sys
path, reps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3
rows=[r for r in csv.reader(open(path)) if len(r)>5]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value"); iu=hdr.index("Metric Unit")
data=rows[1:]
n=len(data)//reps
last=data[(reps-1)*n:]
agg=collections.OrderedDict(); tot=0
for r in last:
    name=r[ik].split('(')[0][-60:]
    v=float(r[iv].replace(',','')); v = v/1000 if r[iu].strip()=='ns' else v
    agg.setdefault(name,[0,0]); agg[name][0]+=v; agg[name][1]+=1; tot+=v
print("launches per step", len(last), "total us %.1f" % tot)
for k,(v,c) in sorted(agg.items(), key=lambda kv:-kv[1][0])[:16]:
    print(f"{v:10.1f} us {100*v/tot:5.1f}%  x{c:4d}  {k}")
