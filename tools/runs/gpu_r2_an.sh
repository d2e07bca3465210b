#!/bin/bash
# round 2, GPU call AN: configs[2] (32 M segments at 512^3) and configs[1] on one GPU with the final kernels: split / frame / packed8, phases; ncu launch list of the split form
mkdir -p gpurun_out
timeout 300 python tools/big_probe.py > gpurun_out/big_probe_an.json 2>gpurun_out/big_probe_an.err; cat gpurun_out/big_probe_an.json | tr -d '\n' | cut -c1-1500; echo
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_red.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_an.csv \
    python tools/big_probe.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_an.csv')))
hdr=None; seen=0
for r in rows:
    if r and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if int(d['ID'])<12: print(d['ID'], d['Kernel Name'][:50], d['Grid Size'], d['Metric Name'], d['Metric Value'])
PY
