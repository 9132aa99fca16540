#!/bin/bash
# usage: profiles/ncu_export.sh <tag> <kernel regex> <skip> <env assignments or ""> [workload]
# Captures ONE launch with `ncu --set full` (48^3 profiling workload by default: a 64^3 handle holds 185 GB and ncu
# cannot save/restore it between replay passes), exports the raw / details / source pages as CSV next to it and
# removes the report (gpurun copies back at most 64 MiB).
tag=$1; re=$2; skip=$3; envs=$4; wl=${5:-cavity3d_48_gh28}
out=gpurun_out/ncu_$tag
env $envs timeout 300 ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c 1 -o $out \
    python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu-baseline > $out.log 2>&1
ncu -i $out.ncu-rep --page raw --csv > $out.raw.csv 2>/dev/null
ncu -i $out.ncu-rep --page details --csv > $out.details.csv 2>/dev/null
ncu -i $out.ncu-rep --page source --csv > $out.source.csv 2>/dev/null
rm -f $out.ncu-rep
ls -la $out.*
