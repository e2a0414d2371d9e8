import csv, sys, collections
fn = sys.argv[1]; nlast = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lines=[l for l in open(fn) if not l.startswith('==')]
seq=[]
for row in csv.DictReader(lines):
    name=row['Kernel Name'].split('(')[0].replace('void ','').replace('prsd::','')
    v=float(row['Metric Value'].replace(',','')); unit=row['Metric Unit']
    if unit=='ns': v/=1e3
    elif unit=='ms': v*=1e3
    seq.append((name,v,row['Grid Size'],row['Block Size']))
# last step = from last k_spline_build(first of step) ... find last 'k_spline_build' with grid (64,1,1)? just take nlast
last=seq[-nlast:]
tot=sum(v for _,v,_,_ in last)
agg=collections.OrderedDict()
for n,v,g,b in last: agg[n]=agg.get(n,0)+v
for n,v,g,b in last: print('%-34s %9.1f us  %s %s'%(n,v,g,b))
print('--- per-kernel share of the step (sum %.1f us)'%tot)
for n,v in sorted(agg.items(), key=lambda x:-x[1]): print('%-34s %9.1f us  %5.1f%%'%(n,v,100*v/tot))
