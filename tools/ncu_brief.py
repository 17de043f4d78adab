import csv,collections,sys
rows=list(csv.DictReader([l for l in open(sys.argv[1]) if l.startswith('"')]))
d=collections.OrderedDict()
for r in rows: d.setdefault((r["ID"],r["Kernel Name"].split('(')[0][-40:],r["Grid Size"]),{})[r["Metric Name"]]=(r["Metric Value"],r["Metric Unit"])
def val(m,n):
    if n not in m: return float('nan')
    v,u=m[n]; v=float(v.replace(',','')); return v*{'ns':1e-3,'nsecond':1e-3,'us':1,'usecond':1,'ms':1e3,'msecond':1e3,'byte':1e-6,'Kbyte':1e-3,'Mbyte':1,'Gbyte':1e3}.get(u,1)
items=list(d.items())
items=items[len(items)//2:] if len(sys.argv)<3 else items
agg=collections.OrderedDict(); tot=0
for k,m in items:
    t=val(m,'gpu__time_duration.sum'); tot+=t
    a=agg.setdefault((k[1],k[2]),[0,0.0,0.0,0.0])
    a[0]+=1; a[1]+=t; a[2]+=val(m,'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'); a[3]+=val(m,'lts__t_bytes.sum')
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print("%-40s %-14s n=%3d total=%9.1f us mean=%8.1f tensor%%=%5.1f L2=%8.1f MB"%(k[0],k[1],a[0],a[1],a[1]/a[0],a[2]/a[0],a[3]/a[0]))
print('total us', round(tot,1))
