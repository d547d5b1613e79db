mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q -x) > gpurun_out/r02g_pytest.log 2>&1; tail -5 gpurun_out/r02g_pytest.log
for cfg in "16 64" "0 64" "16 32" "0 32"; do
  set -- $cfg
  X265CU_GREEN=$1 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-others --async-depth $2 > gpurun_out/r02g_g$1_a$2.json 2> gpurun_out/r02g_g$1_a$2.err || tail -5 gpurun_out/r02g_g$1_a$2.err
  python -c "
import json
d=json.load(open('gpurun_out/r02g_g$1_a$2.json')); r=d['roofline']
print('green $1 async $2: value %.1f %s e2e %.1f %s us/job %.1f' % (d['value'], d['ms_steps'], d['e2e']['value'], d['e2e']['ms_steps'], r['search_us_per_job'])); print('  kern', r['kernel_busy_ms_per_step']); print('  host', r['host_ms_per_step']); print('  e2e ', d['e2e']['host_ms_last_step'])"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lowres_fused -s 20 -c 1 -o gpurun_out/prof_lowres_r02b python bench.py --no-cpu-baseline --no-others --no-e2e --steps 1 --warmup 0 --frames 60 > gpurun_out/ncu_lowres.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:extend_border -s 20 -c 1 -o gpurun_out/prof_border_r02b python bench.py --no-cpu-baseline --no-others --no-e2e --steps 1 --warmup 0 --frames 60 > gpurun_out/ncu_border.log 2>&1
