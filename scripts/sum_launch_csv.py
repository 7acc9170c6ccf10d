import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
tot=0
from collections import OrderedDict
agg=OrderedDict()
for r in rows[1:]:
    k=r[ki][:58]; v=float(r[vi].replace(",",""))/1e3
    agg.setdefault(k,[]).append(v); tot+=v
for k,v in agg.items(): print("%-60s %s" % (k, " ".join("%.1f"%x for x in v)))
print("total us", round(tot,1))
