# Round-2 ncu captures (one gpurun call; every command has run without ncu first)
set -x
for w in c2 t128 c3 c4; do python scripts/ncu_targets.py $w > gpurun_out/r02_plain_$w.log 2>&1 || exit 1; done
M=gpu__time_duration.sum
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_c2.csv python scripts/ncu_targets.py c2 > /dev/null 2>&1
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_c3.csv python scripts/ncu_targets.py c3 > /dev/null 2>&1
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_c4.csv python scripts/ncu_targets.py c4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_filter -c 2 -o gpurun_out/r02_full_c2 python scripts/ncu_targets.py c2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_filter -c 1 -o gpurun_out/r02_full_t128 python scripts/ncu_targets.py t128 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_filter -c 1 -o gpurun_out/r02_full_c3 python scripts/ncu_targets.py c3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:radius_kernel -c 1 -o gpurun_out/r02_full_c4 python scripts/ncu_targets.py c4 > /dev/null 2>&1
# gpurun_out is limited to 64 MiB: keep the C2 report for the source page, the raw metric pages of the others
for w in c2 t128 c3 c4; do ncu -i gpurun_out/r02_full_$w.ncu-rep --page raw --csv > gpurun_out/r02_full_${w}_raw.csv 2>/dev/null; done
rm -f gpurun_out/r02_full_t128.ncu-rep gpurun_out/r02_full_c3.ncu-rep gpurun_out/r02_full_c4.ncu-rep
ls -la gpurun_out/r02_*
