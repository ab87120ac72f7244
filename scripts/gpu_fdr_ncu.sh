# ncu of the FDR micro-benchmark variants given as arguments (scripts/ubin/<name>): pipe utilisation, issue rate, stall reasons of the loop alone
set -x
mkdir -p gpurun_out /tmp/ncu
for n in "$@"; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:fdr_bench -s 2 -c 1 -o /tmp/ncu/fdr_$n -f scripts/ubin/$n > /tmp/ncu/fdr_$n.log 2>&1; tail -1 /tmp/ncu/fdr_$n.log
  python scripts/ncu_summary.py /tmp/ncu/fdr_$n.ncu-rep > gpurun_out/ncu_fdr_$n.txt 2>&1
  ncu -i /tmp/ncu/fdr_$n.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[2]
for i,n in enumerate(h):
    if any(k in n for k in ('pipe_','inst_executed_pipe','issue_slot','warp_issue_stalled','dispatch','reg_bank','smsp__inst_issued','smsp__warps_eligible','register')): print(f'{n:90s} {v[i]}')
" >> gpurun_out/ncu_fdr_$n.txt
  ncu -i /tmp/ncu/fdr_$n.ncu-rep --page source --csv --print-source sass > /tmp/ncu/src_$n.csv 2>/dev/null
  python - $n <<'PY'
import csv,sys
n=sys.argv[1]
rows=list(csv.reader(open(f'/tmp/ncu/src_{n}.csv')))
hdr=None
out=open(f'gpurun_out/ncu_fdr_{n}_sass.txt','w')
for r in rows:
    if r and r[0] in ('Address','#'):
        hdr=r; idx={c:i for i,c in enumerate(hdr)}; print('columns:',[c for c in hdr][:40],file=out); continue
    if hdr is None or len(r)<len(hdr): continue
    try:
        inst=int(r[idx['Instructions Executed']] or 0)
    except Exception: continue
    if inst < 100000: continue
    keep=[c for c in hdr if c.startswith('stall_') or c in ('# Samples','Instructions Executed','Source','Warp Stall Sampling (All Samples)','Warp Stall Sampling (Not-issued Samples)')]
    print(' | '.join(f'{r[idx[c]]}' if c=='Source' else f'{c[:14]}={r[idx[c]]}' for c in keep if r[idx[c]] not in ('0','')),file=out)
PY
done
