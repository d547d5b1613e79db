mkdir -p gpurun_out
(time timeout 300 python bench.py --steps 5 --warmup 3 --no-others --no-cpu-baseline) > gpurun_out/r02w_bench_e2e.json 2> gpurun_out/r02w_bench_e2e.err; tail -3 gpurun_out/r02w_bench_e2e.err
python - <<'PY'
import json
for l in open("gpurun_out/r02w_bench_e2e.json"):
    l = l.strip()
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], d["ms_steps"], "e2e", d["e2e"]["value"], d["e2e"]["ms_steps"], d["e2e"]["host_ms_last_step"])
PY
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv
