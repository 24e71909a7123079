#!/bin/bash
# ncu captures of one round (run under gpurun, one GPU): launch list of synchronous steps + full captures of the top kernels.
#   bash scripts/ncu_round.sh r02j "knn_mma_kernel orb_level_kernel lk_track2_kernel"
tag=${1:-rXX}
kernels=${2:-"lk_track2_kernel"}
out=gpurun_out
python scripts/profile_step.py 32 4 > $out/${tag}_step_plain.txt 2>&1 || { echo "plain run failed"; tail -5 $out/${tag}_step_plain.txt; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/${tag}_launches_step.csv \
    python scripts/profile_step.py 32 4 > $out/${tag}_ncu_launches.log 2>&1
for k in $kernels; do
  ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip ${SKIP:-2} -c ${COUNT:-2} -o $out/${tag}_$k -f \
      python scripts/profile_step.py 32 2 > $out/${tag}_ncu_$k.log 2>&1
  ncu -i $out/${tag}_$k.ncu-rep --page details > $out/${tag}_${k}_details.txt 2>/dev/null
  ncu -i $out/${tag}_$k.ncu-rep --page raw --csv > $out/${tag}_${k}_raw.csv 2>/dev/null
  ncu -i $out/${tag}_$k.ncu-rep --page source --print-source cuda,sass --csv > $out/${tag}_${k}_source.csv 2>/dev/null
done
ls -la $out | tail -20
