#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:"k1_" --csv --log-file gpurun_out/q_traffic.csv \
  python tools/ab_box.py --nb 64 --steps 1 --warmup 1 --repeat 1 --profile-steps 0 "m|fast|" "m32|fast|l2_fetch=32" "m128|fast|l2_fetch=128"  "xs16_32|fast|block_order=xslab16,l2_fetch=32" "xs16_64|fast|block_order=xslab16,l2_fetch=64" "xs12_64|fast|block_order=xslab12,l2_fetch=64" "xs20_64|fast|block_order=xslab20,l2_fetch=64" "xs24_64|fast|block_order=xslab24,l2_fetch=64" > gpurun_out/q_ncu.log 2>&1; echo "exit $?" >> gpurun_out/q_ncu.log
tail -3 gpurun_out/q_ncu.log | cut -c1-200
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/q_traffic.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]; h=rows[hi]
d=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<len(h): continue
    x=dict(zip(h,r)); d.setdefault((int(x['ID']),x['Kernel Name'][:50],x['Grid Size']),{})[x['Metric Name']]=float(x['Metric Value'].replace(',',''))
for k,m in d.items():
    if '(8192' in k[2] or '(32768' in k[2] or '(16384' in k[2]: continue
    print(k, 't=%.3fms rd=%.3fGB wr=%.3fGB L2hit=%.1f L1hit=%.1f BW=%.2fTB/s'%(m['gpu__time_duration.sum']/1e6,m['dram__bytes_read.sum']/1e9,m['dram__bytes_write.sum']/1e9,m['lts__t_sector_hit_rate.pct'],m['l1tex__t_sector_hit_rate.pct'],(m['dram__bytes_read.sum']+m['dram__bytes_write.sum'])/m['gpu__time_duration.sum']/1e3))
PY
