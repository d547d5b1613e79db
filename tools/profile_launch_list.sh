# per-launch durations (serialised by ncu) of the first 2000 launches of one short bench step
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r01m.csv python bench.py --no-cpu-baseline --steps 1 --warmup 0 --frames 110 > gpurun_out/ncu_launches.log 2>&1
tail -n 1 gpurun_out/ncu_launches.log | cut -c1-200
