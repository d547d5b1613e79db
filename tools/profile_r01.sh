# full ncu captures of the two hot kernels of one short bench step (first launches = the saturated ones)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_kernel -c 1 -o gpurun_out/prof_search_r01k python bench.py --no-cpu-baseline --steps 1 --warmup 0 --frames 110 > gpurun_out/ncu_search.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_group_kernel -c 1 -o gpurun_out/prof_cost_r01k python bench.py --no-cpu-baseline --steps 1 --warmup 0 --frames 110 > gpurun_out/ncu_cost.log 2>&1
tail -n 2 gpurun_out/ncu_search.log; tail -n 2 gpurun_out/ncu_cost.log
for d in 48 40; do
timeout 300 python bench.py --no-cpu-baseline --async-depth $d > gpurun_out/sd_$d.json 2> gpurun_out/sd_$d.err
done
