#!/bin/bash
# Round profile collection on the GPU box (run through gpurun from the repo root):
#   bash scripts/collect_profiles.sh <tag>      e.g. r01
# 1. plain bench (no profiler) -> gpurun_out/<tag>_bench.json
# 2. ncu launch list of the same command -> gpurun_out/<tag>_launches.csv + summary
# 3. ncu --set full of the top kernels (few launches each, no source import) -> summaries; reports deleted if large
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
rm -f $out/*.ncu-rep
set -x
python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err || { tail -20 $out/${tag}_bench.err; exit 1; }
tail -c 600 $out/${tag}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --slab-grid 0 --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1
python scripts/summarize_launches.py $out/${tag}_launches.csv --tail 2000 > $out/${tag}_launches_summary.txt
head -25 $out/${tag}_launches_summary.txt
# one step's worth of launches of each kernel, taken after the three warm-up steps (8 / 4 / 4 launches per step)
for spec in k_fft_cols:24:8 k_cols_mult:12:4 k_gemv:12:4; do
  k=${spec%%:*}; r=${spec#*:}; skip=${r%%:*}; cnt=${r#*:}
  ncu --set full --clock-control none -k regex:^${k} -s $skip -c $cnt -f -o $out/${tag}_${k} \
      python bench.py --steps 1 --warmup 3 --slab-grid 0 --no-cpu-baseline > $out/${tag}_ncu_${k}.log 2>&1
  if [ -f $out/${tag}_${k}.ncu-rep ]; then
    python scripts/ncu_summary.py $out/${tag}_${k}.ncu-rep $out/${tag}_ncu_${k} > /dev/null
    sz=$(stat -c %s $out/${tag}_${k}.ncu-rep)
    if [ $sz -gt 8000000 ]; then rm -f $out/${tag}_${k}.ncu-rep; fi
  fi
done
du -sh $out
# Stokes solve timing (rows a5, a6, a10, a12) and the slab grid solve kernel split
python scripts/bench_stokes.py --grid 1024 --check --profile > $out/${tag}_stokes_1024.json 2> $out/${tag}_stokes_1024.err
tail -c 400 $out/${tag}_stokes_1024.json
python scripts/run_slab_multi.py --grid 16384 --profile > $out/${tag}_slab_16384_1gpu.txt 2>&1
cat $out/${tag}_slab_16384_1gpu.txt | tail -6
