mkdir -p gpurun_out
python tools/h2d_probe.py 2>&1 | tail -2
(time timeout 900 python -m pytest tests -m gpu -q -x) > gpurun_out/r02f_pytest.log 2>&1; tail -5 gpurun_out/r02f_pytest.log
for cfg in "16 64" "0 64" "16 32" "16 96"; do
  set -- $cfg
  X265CU_GREEN=$1 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-others --async-depth $2 > gpurun_out/r02f_g$1_a$2.json 2> gpurun_out/r02f_g$1_a$2.err || tail -5 gpurun_out/r02f_g$1_a$2.err
  python -c "
import json
d=json.load(open('gpurun_out/r02f_g$1_a$2.json')); r=d['roofline']
print('green $1 async $2: value %.1f e2e %.1f us/job %.1f part %s' % (d['value'], d['e2e']['value'], r['search_us_per_job'], d['config'].get('sm_partition'))); print('  kern', r['kernel_busy_ms_per_step']); print('  host', r['host_ms_per_step']); print('  e2e ', d['e2e']['host_ms_last_step'])"
done
