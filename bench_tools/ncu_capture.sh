#!/bin/bash
# ncu_capture.sh <tag> <kernel-regex> <command...>: one `ncu --set full` capture of the first matching launch after 4
# skipped ones, exported as text next to the report (gpurun_out/ has a size cap: the report itself is dropped when large).
#   gpurun_out/<tag>_details.csv   --page details (sections: SOL, memory, scheduler, warp states, occupancy ...)
#   gpurun_out/<tag>_source.csv.gz --page source  (SASS with per-instruction samples / stall reasons)
tag=$1; shift
kre=$1; shift
out=gpurun_out/$tag
ncu --set full --import-source on --clock-control none -k "regex:$kre" --launch-skip 4 --launch-count 1 -o $out -f "$@" > ${out}_ncu.log 2>&1
tail -2 ${out}_ncu.log
ncu -i $out.ncu-rep --page details --csv > ${out}_details.csv 2>/dev/null
ncu -i $out.ncu-rep --page raw --csv > ${out}_raw.csv 2>/dev/null
ncu -i $out.ncu-rep --page source --csv 2>/dev/null | gzip -9 > ${out}_source.csv.gz
sz=$(stat -c %s $out.ncu-rep)
if [ "$sz" -gt 20000000 ]; then rm -f $out.ncu-rep; fi
ls -la gpurun_out/ | grep $tag
