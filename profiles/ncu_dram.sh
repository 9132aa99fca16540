#!/bin/bash
# usage: ncu_dram.sh tag "ENV=.." kernel-regex
tag=$1; envs=$2; re=$3
env $envs timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:$re -s 30 -c 6 --csv --log-file gpurun_out/dram_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/dram_$tag.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/dram_$tag.csv')) if len(r)>10]
hdr=rows[0]; iN=hdr.index('Kernel Name'); iM=hdr.index('Metric Name'); iV=hdr.index('Metric Value'); iI=hdr.index('ID')
d={}
for r in rows[1:]:
    d.setdefault(r[iI],{})[r[iM]]=float(r[iV].replace(',','')); d[r[iI]]['name']=r[iN][:40]
for k,v in d.items(): print('$tag',k,v)
PY
