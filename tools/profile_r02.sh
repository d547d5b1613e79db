# round-2 evidence run on the GPU box: timelines, residency-cap sweep, ncu launch list and full captures
mkdir -p gpurun_out
rm -f gpurun_out/tl_value.csv gpurun_out/tl_e2e.csv
X265CU_TIMELINE=gpurun_out/tl_value.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-others > gpurun_out/tl_value.json 2> gpurun_out/tl_value.err
python tools/timeline_summary.py gpurun_out/tl_value.csv 100 -1 > gpurun_out/tl_value.txt; cat gpurun_out/tl_value.txt
X265CU_TIMELINE=gpurun_out/tl_e2e.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-others > gpurun_out/tl_e2e.json 2> gpurun_out/tl_e2e.err
python tools/timeline_summary.py gpurun_out/tl_e2e.csv 100 -1 > gpurun_out/tl_e2e.txt; cat gpurun_out/tl_e2e.txt
python -c "
import json
d=json.load(open('gpurun_out/tl_e2e.json')); print('value %.1f e2e %.1f' % (d['value'], d['e2e']['value'])); print(d['e2e']['host_ms_last_step'])"
for sm in 9216 10240 12288; do
  X265CU_SEARCH_SMEM=$sm python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-others > gpurun_out/smem_$sm.json 2> gpurun_out/smem_$sm.err
  python -c "
import json
d=json.load(open('gpurun_out/smem_$sm.json')); r=d['roofline']; print('smem $sm: value %.1f e2e %.1f us/job %.1f' % (d['value'], d['e2e']['value'], r['search_us_per_job']), d['e2e']['host_ms_last_step'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --no-cpu-baseline --no-others --no-e2e --steps 1 --warmup 0 --frames 120 > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lowres_fused -s 20 -c 1 -o gpurun_out/prof_lowres_r02 python bench.py --no-cpu-baseline --no-others --no-e2e --steps 1 --warmup 0 --frames 60 > gpurun_out/ncu_lowres.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 4 -c 1 -o gpurun_out/prof_search_r02 python bench.py --no-cpu-baseline --no-others --no-e2e --steps 1 --warmup 0 --frames 200 > gpurun_out/ncu_search.log 2>&1
ls -la gpurun_out/*.ncu-rep
