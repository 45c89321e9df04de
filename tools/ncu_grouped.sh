#!/bin/bash
# usage (on the GPU box): tools/ncu_grouped.sh <tag>   -> gpurun_out/r2/<tag>.ncu-rep + raw/source CSV exports
# one `ncu --set full` capture of k_metrics_grouped on the 4M-polyline bench (after the same command ran clean without ncu)
tag=$1; shift
mkdir -p gpurun_out/r2
python bench.py --steps 2 --warmup 3 --streamlines 4000000 --no-cpu-baseline "$@" > gpurun_out/r2/${tag}_plain.json 2> gpurun_out/r2/${tag}_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_metrics_grouped -s 3 -c 1 -f -o gpurun_out/r2/${tag} \
    python bench.py --steps 2 --warmup 3 --streamlines 4000000 --no-cpu-baseline "$@" > gpurun_out/r2/${tag}_ncu.log 2>&1
ncu -i gpurun_out/r2/${tag}.ncu-rep --page raw --csv > gpurun_out/r2/${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/r2/${tag}.ncu-rep --page source --csv > gpurun_out/r2/${tag}_src.csv 2>/dev/null
