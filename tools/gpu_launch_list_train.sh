#!/bin/bash
# ncu launch list of the training step; prints the kernels of the LAST step (after the two warm-up steps) outside csrc/train.cu
O=gpurun_out; mkdir -p $O
timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/l1_launches_train.csv python tools/train_step_target.py 1 0 8 > $O/l1_ncu.log 2>&1
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/l1_launches_train.csv') if l.startswith('"')]
rd=csv.reader(lines); hdr=next(rd)
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
data=[(r[ki], float(r[vi].replace(',',''))) for r in rd]
idx=[i for i,(k,v) in enumerate(data) if 'pack_weights' in k]
last=data[idx[-1]:]
tot=collections.Counter(); cnt=collections.Counter()
for k,v in last:
    name=k.split('(')[0][:78]
    if 'mz::' in k: name='MZ:'+k.split('::')[-1].split('(')[0][:40]
    tot[name]+=v/1e3; cnt[name]+=1
print('kernels in the last step:', len(last))
for n,t in tot.most_common(60): print('%8.1f us %4d  %s'%(t,cnt[n],n))
print('non-mz total us %.1f in %d launches' % (sum(t for n,t in tot.items() if not n.startswith('MZ:')), sum(c for n,c in cnt.items() if not n.startswith('MZ:'))))
PY
