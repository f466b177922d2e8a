#!/bin/bash
# Reduced version of collect_profiles.sh after a change to one kernel: bench, launch list, ncu --set full of $2 only
tag=${1:-r01}; k=${2:-k_cols_mult}; skip=${3:-12}; cnt=${4:-4}
out=gpurun_out
mkdir -p $out
python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err || { tail -20 $out/${tag}_bench.err; exit 1; }
tail -c 300 $out/${tag}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --slab-grid 0 --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none -k regex:^${k} -s $skip -c $cnt -f -o $out/${tag}_${k} \
    python bench.py --steps 1 --warmup 3 --slab-grid 0 --no-cpu-baseline > $out/${tag}_ncu_${k}.log 2>&1
python scripts/ncu_summary.py $out/${tag}_${k}.ncu-rep $out/${tag}_ncu_${k} > /dev/null
rm -f $out/${tag}_${k}.ncu-rep
