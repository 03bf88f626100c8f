# Round-2 captures of config 3 with the two-means partition (the command ran without ncu first)
set -x
python scripts/ncu_targets.py c3 > gpurun_out/r02_plain_c3p.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c3p.csv python scripts/ncu_targets.py c3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ball_tile_kernel|knn_filter_kernel|seed_bound_kernel" -c 4 -o gpurun_out/r02_full_c3p python scripts/ncu_targets.py c3 > /dev/null 2>&1
ncu -i gpurun_out/r02_full_c3p.ncu-rep --page raw --csv > gpurun_out/r02_full_c3p_raw.csv 2>/dev/null
rm -f gpurun_out/r02_full_c3p.ncu-rep
ls -la gpurun_out/r02_*c3p*
