#!/bin/bash
# round-2 profile collection (run through gpurun): launch list of the default bench command and ncu --set full of the
# dominant kernel, summarised by scripts/ncu_summary.py.  $1 = kernel regex (default k_fft_cols)
k=${1:-k_fft_cols}
out=gpurun_out
mkdir -p $out
args="--steps 2 --warmup 3 --slab-grid 0 --no-cpu-baseline --stokes-grid 0"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file $out/r02_launches.csv \
    python bench.py $args > $out/r02_ncu_launches.log 2>&1
python scripts/summarize_launches.py $out/r02_launches.csv > $out/r02_launches_summary.txt 2>&1 || true
ncu --set full --clock-control none --import-source on -k regex:^${k} -s 21 -c 7 -f -o $out/r02_${k} \
    python bench.py $args > $out/r02_ncu_${k}.log 2>&1
python scripts/ncu_summary.py $out/r02_${k}.ncu-rep $out/r02_ncu_${k} > /dev/null
ls -la $out/r02_*
