# launch list of one short bench step (per-launch durations, serialised by ncu) + full captures of the two hot kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r01h.csv python bench.py --no-cpu-baseline --steps 1 --warmup 0 --frames 90 > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 3 -c 1 -o gpurun_out/prof_search_r01h python bench.py --no-cpu-baseline --steps 1 --warmup 0 --frames 90 > gpurun_out/ncu_search.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_group_kernel -s 3 -c 1 -o gpurun_out/prof_cost_r01h python bench.py --no-cpu-baseline --steps 1 --warmup 0 --frames 90 > gpurun_out/ncu_cost.log 2>&1
tail -2 gpurun_out/ncu_launches.log gpurun_out/ncu_search.log gpurun_out/ncu_cost.log
