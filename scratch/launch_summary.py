import csv, collections, re, sys
path=sys.argv[1]
with open(path) as f:
    lines=[l for l in f if not l.startswith('==')]
r=csv.DictReader(lines)
agg=collections.OrderedDict()
for row in r:
    name=re.sub(r'\(.*','',row['Kernel Name'])[:72]
    try: v=float(row['Metric Value'].replace(',',''))
    except: continue
    unit=row['Metric Unit']
    if unit in ('ns','nsecond'): v/=1000
    elif unit in ('ms','msecond'): v*=1000
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
print('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)')
print('total_us %.1f over %d launches'%(tot,sum(a[0] for a in agg.values())))
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print('%-74s n=%4d sum_us=%9.1f avg_us=%8.2f share=%5.1f%%'%(k,c,t,t/c,100*t/tot))
