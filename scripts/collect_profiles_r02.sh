#!/bin/bash
# round-2 profile collection (run through gpurun).  The launch list and the ncu --set full capture use --config l
# (8192^2: the same kernels and launch sequence as the 16384^2 default at a quarter of the ncu replay time).
# $1 = kernel regex of the full capture (default k_cols_mult), $2 = launches to skip, $3 = launches to capture
k=${1:-k_cols_mult}; skip=${2:-9}; cnt=${3:-6}
out=gpurun_out
mkdir -p $out
args="--config l --steps 2 --warmup 3 --slab-grid 0 --no-cpu-baseline --stokes-grid 0"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file $out/r02_launches.csv \
    python bench.py $args > $out/r02_ncu_launches.log 2>&1
python scripts/summarize_launches.py $out/r02_launches.csv > $out/r02_launches_summary.txt 2>&1 || true
ncu --set full --clock-control none --import-source on -k regex:^${k} -s $skip -c $cnt -f -o $out/r02_${k} \
    python bench.py $args > $out/r02_ncu_${k}.log 2>&1
python scripts/ncu_summary.py $out/r02_${k}.ncu-rep $out/r02_ncu_${k} > /dev/null
ls -la $out/r02_*
