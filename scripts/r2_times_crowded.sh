#!/bin/bash
# per-launch durations of one fit batch (default 64 FFIs): ncu time-only pass
mkdir -p gpurun_out
N=${1:-64}
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct --clock-control none -k regex:"^k_" -s 33 -c 33 --csv --log-file gpurun_out/times.csv python scripts/prof_run_crowded.py $N > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/times.csv')) if len(r)>10]
hdr=rows[0]; i_n=hdr.index('Kernel Name'); i_m=hdr.index('Metric Name'); i_v=hdr.index('Metric Value'); i_id=hdr.index('ID'); i_u=hdr.index('Metric Unit')
d={}
def f(x):
    try: return float(x.replace(',',''))
    except: return 0.0
for r in rows[1:]:
    v=f(r[i_v])
    if r[i_m]=='gpu__time_duration.sum' and r[i_u]=='ns': v/=1000
    if r[i_m]=='gpu__time_duration.sum' and r[i_u]=='ms': v*=1000
    d.setdefault((r[i_id],r[i_n].split('(')[0]),{})[r[i_m]]=v
tot=0
for (id_,name),m in d.items():
    tot+=m.get('gpu__time_duration.sum',0)
    print(f"{name[:40]:40s} {m.get('gpu__time_duration.sum',0):9.1f} us  inst {m.get('smsp__inst_executed.sum',0)/1e6:7.1f} M  warps {m.get('sm__warps_active.avg.pct_of_peak_sustained_active',0):5.1f}%")
print('total us', tot)
PY
