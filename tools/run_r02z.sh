mkdir -p gpurun_out
(time timeout 200 python bench.py --no-others --no-cpu-baseline --steps 2 --warmup 2 --frames 150) > gpurun_out/r02z_bench_quick.json 2> gpurun_out/r02z_bench_quick.err; tail -3 gpurun_out/r02z_bench_quick.err
python - <<'PY'
import json
for l in open("gpurun_out/r02z_bench_quick.json"):
    l = l.strip()
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], d["ms_steps"], "e2e", d["e2e"]["value"], d["e2e"]["ms_steps"])
PY
