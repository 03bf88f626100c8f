# Round-2 ncu captures of the warp-per-query scan (one gpurun call; the commands run without ncu first)
set -x
for w in c4knn c1; do python scripts/ncu_targets.py $w > gpurun_out/r02_plain_$w.log 2>&1 || exit 1; done
M=gpu__time_duration.sum
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_c4knn.csv python scripts/ncu_targets.py c4knn > /dev/null 2>&1
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_c1.csv python scripts/ncu_targets.py c1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_warp -c 1 -o gpurun_out/r02_full_c4knn python scripts/ncu_targets.py c4knn > /dev/null 2>&1
ncu -i gpurun_out/r02_full_c4knn.ncu-rep --page raw --csv > gpurun_out/r02_full_c4knn_raw.csv 2>/dev/null
ls -la gpurun_out/r02_*c4knn* gpurun_out/r02_*c1*
