# Round bookkeeping run on one B200: GPU tests, default bench, launch list + full ncu capture of the filter kernel, T / C3 configs.
set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/r01b_bench.json 2> gpurun_out/r01b_bench.err; tail -c 1500 gpurun_out/r01b_bench.json
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --queries 151552 > gpurun_out/r01b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches_tc.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --queries 151552 > gpurun_out/r01b_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_filter -s 1 -c 1 -o gpurun_out/r01b_filter python bench.py --steps 2 --warmup 1 --no-cpu-baseline --queries 151552 > gpurun_out/r01b_ncu_f.log 2>&1
python scripts/run_configs.py ${CONFIGS:-t128 c3} > gpurun_out/r01b_configs.jsonl 2> gpurun_out/r01b_configs.err; cat gpurun_out/r01b_configs.jsonl | cut -c1-700
