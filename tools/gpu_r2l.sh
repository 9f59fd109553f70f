#!/bin/bash
mkdir -p gpurun_out
for cfg in 1,16 1,8 2,8 2,6 2,4 3,5 3,4 4,4; do
  echo "cfg $cfg" >> gpurun_out/r2l_explore.txt
  B200_TS_EXPLORE=$cfg ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bruss3d" -c 60 --csv --log-file gpurun_out/r2l_tmp.csv python tools/stencil_bench.py 100 > /dev/null 2>&1
  python - <<'PY' >> gpurun_out/r2l_explore.txt
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2l_tmp.csv')))
hdr=None; agg=collections.defaultdict(list)
for r in rows:
    if r and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); agg[d['Kernel Name'][18:45]+d['Grid Size']].append(float(d['Metric Value'].replace(',','')))
for k,v in agg.items(): print('  ',k, round(sum(v)/len(v)), round(min(v)), len(v))
PY
done
cat gpurun_out/r2l_explore.txt
