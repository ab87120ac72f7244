set -x
mkdir -p gpurun_out /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lde_step -s 60 -c 1 -o /tmp/ncu/lde -f python bench.py --workload lde --steps 1 --warmup 0 --runs 256 --no-cpu-baseline --stripes 1 --no-graph > /tmp/ncu/lde.log 2>&1; tail -2 /tmp/ncu/lde.log
python scripts/ncu_summary.py /tmp/ncu/lde.ncu-rep > gpurun_out/r2zz_ncu_lde_step_d30.txt 2>&1
ncu -i /tmp/ncu/lde.ncu-rep --page source --csv --print-source sass,cuda > /tmp/ncu/src.csv 2>/dev/null
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('/tmp/ncu/src.csv')))
cur=None; hdr=None
agg=collections.defaultdict(lambda:[0,0]); filetot=collections.defaultdict(lambda:[0,0])
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur=r[1].split('/')[-1]; hdr=None; continue
    if r and r[0]=='Line No': hdr=r; idx={n:i for i,n in enumerate(hdr)}; continue
    if hdr is None or not r or not r[0].isdigit(): continue
    try: inst=int(r[idx['Instructions Executed']] or 0); smp=int(r[idx['# Samples']] or 0)
    except: continue
    agg[(cur,int(r[0]),r[1][:110])][0]+=inst; agg[(cur,int(r[0]),r[1][:110])][1]+=smp
    filetot[cur][0]+=inst; filetot[cur][1]+=smp
tot=sum(v[0] for v in filetot.values()); ts=sum(v[1] for v in filetot.values())
out=open('gpurun_out/r2zz_ncu_lde_hot_lines.txt','w')
print('total warp instructions',tot,'stall samples',ts,file=out)
for f,v in sorted(filetot.items(), key=lambda kv:-kv[1][0])[:10]: print(f'{f:28s} inst {100*v[0]/tot:5.1f}%  samples {100*v[1]/ts:5.1f}%',file=out)
print(file=out)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:70]: print(f'{k[0]:22s}:{k[1]:4d} inst {100*v[0]/tot:5.2f}% smp {100*v[1]/ts:5.2f}%  {k[2]}',file=out)
PY
mkdir -p gpurun_out /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gleet_step -s 60 -c 1 -o /tmp/ncu/gls -f python bench.py --workload gleet --steps 1 --warmup 0 --runs 256 --no-cpu-baseline --stripes 1 --no-graph > /tmp/ncu/gls.log 2>&1; tail -2 /tmp/ncu/gls.log
python scripts/ncu_summary.py /tmp/ncu/gls.ncu-rep > gpurun_out/r2zz_ncu_gleet_step.txt 2>&1
ncu -i /tmp/ncu/gls.ncu-rep --page source --csv --print-source sass,cuda > /tmp/ncu/src.csv 2>/dev/null
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('/tmp/ncu/src.csv')))
cur=None; hdr=None
agg=collections.defaultdict(lambda:[0,0]); filetot=collections.defaultdict(lambda:[0,0])
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur=r[1].split('/')[-1]; hdr=None; continue
    if r and r[0]=='Line No': hdr=r; idx={n:i for i,n in enumerate(hdr)}; continue
    if hdr is None or not r or not r[0].isdigit(): continue
    try: inst=int(r[idx['Instructions Executed']] or 0); smp=int(r[idx['# Samples']] or 0)
    except: continue
    agg[(cur,int(r[0]),r[1][:110])][0]+=inst; agg[(cur,int(r[0]),r[1][:110])][1]+=smp
    filetot[cur][0]+=inst; filetot[cur][1]+=smp
tot=sum(v[0] for v in filetot.values()); ts=sum(v[1] for v in filetot.values())
out=open('gpurun_out/r2zz_ncu_gleet_step_hot_lines.txt','w')
print('total warp instructions',tot,'stall samples',ts,file=out)
for f,v in sorted(filetot.items(), key=lambda kv:-kv[1][0])[:10]: print(f'{f:28s} inst {100*v[0]/tot:5.1f}%  samples {100*v[1]/ts:5.1f}%',file=out)
print(file=out)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:70]: print(f'{k[0]:22s}:{k[1]:4d} inst {100*v[0]/tot:5.2f}% smp {100*v[1]/ts:5.2f}%  {k[2]}',file=out)
PY
