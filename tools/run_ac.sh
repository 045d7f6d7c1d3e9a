#!/bin/bash
# ncu --set full over every conv-stack launch of one cfg-2 step (final code of the round)
mkdir -p gpurun_out
timeout -s KILL 1500 ncu --set full --import-source on --clock-control none -k regex:"conv_igemm|conv_pair|maxpool|pack_image" -s 244 -c 62 -o /tmp/cfg2_full python bench.py --config cfg2 --steps 2 --warmup 3 --no-cpu --no-selfcheck --no-sections > /dev/null 2>&1
ncu -i /tmp/cfg2_full.ncu-rep --page raw --csv > gpurun_out/r2ac_ncu_full_cfg2_conv_step.csv 2>/dev/null
ls -la gpurun_out/r2ac* /tmp/cfg2_full.ncu-rep
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2ac_ncu_full_cfg2_conv_step.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]
def col(n): return h.index(n)
tot=0
for r in rows[hi+2:]:
    if len(r)<=col('gpu__time_duration.sum'): continue
    t=float(r[col('gpu__time_duration.sum')]); tot+=t
    print(r[col('Kernel Name')][:44].replace('void ',''), r[col('Grid Size')], round(t,1), 'us  tensor', r[col('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active')] if 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active' in h else '', 'dram%', r[col('dram__throughput.avg.pct_of_peak_sustained_elapsed')][:5] if 'dram__throughput.avg.pct_of_peak_sustained_elapsed' in h else '')
print('total us', tot)
PY
