mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_trainer.py -q -x --timeout 400 --timeout-method=thread 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fused_train_step.csv python tools/fused_train_steps.py 3 > gpurun_out/fts.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_fused_train_step.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn,mv=h.index('Kernel Name'),h.index('Metric Value')
names=[(r[kn].split('(')[0].replace('void ','')[:44], float(r[mv].replace(',',''))) for r in rows[hi+2:]]
start=[i for i,(n,_) in enumerate(names) if 'pack_rays' in n][-1]
agg=collections.OrderedDict()
for n,v in names[start:]:
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
print(len(names[start:]), 'launches', tot/1e6, 'ms')
for n,(c,v) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f'{n:46s} x{c:2d} {v/1e3:8.1f} us')
PY
