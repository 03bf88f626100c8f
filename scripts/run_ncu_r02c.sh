# Round-2 launch lists with the final library (the commands ran without ncu in run_ncu_r02.sh / r02b)
set -x
M=gpu__time_duration.sum
for w in c2 c3 c4; do
  python scripts/ncu_targets.py $w > gpurun_out/r02_plain_$w.log 2>&1 || exit 1
  ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_$w.csv python scripts/ncu_targets.py $w > /dev/null 2>&1
done
ls -la gpurun_out/r02_launches_*.csv
