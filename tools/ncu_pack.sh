#!/bin/bash
# run on the GPU box after a capture: summarise gpurun_out/NAME.ncu-rep into gpurun_out/NAME.md, keep the raw and source pages
# as gzip'd CSV and drop the report itself (gpurun copies back at most 64 MiB)
#   tools/ncu_pack.sh NAME "title" [warps frames]
N=$1; T=$2; shift 2
python tools/ncu_summary.py gpurun_out/$N.ncu-rep gpurun_out/$N.md "$T" "$@" > /dev/null 2>&1
ncu -i gpurun_out/$N.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/$N.raw.csv.gz
ncu -i gpurun_out/$N.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/$N.source.csv.gz
rm -f gpurun_out/$N.ncu-rep
