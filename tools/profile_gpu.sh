#!/bin/bash
# GPU box: one `ncu --set full` capture of a fused kernel + the raw/source pages as CSV (read here with profiles/summarize.py).
# usage: tools/profile_gpu.sh <tag> <kernel regex> <launch skip> <bench.py workload args...>
tag=$1; kre=$2; skip=$3; shift 3
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 -f -o gpurun_out/prof_$tag \
    python bench.py "$@" --steps 1 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/raw_$tag.csv 2>/dev/null
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv > gpurun_out/src_$tag.csv 2>/dev/null
rm -f gpurun_out/prof_$tag.ncu-rep      # gpurun_out/ comes back only if it stays under 64 MiB: the CSV pages are what is read
ls -l gpurun_out/raw_$tag.csv gpurun_out/src_$tag.csv
