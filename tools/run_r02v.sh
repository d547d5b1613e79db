mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/r02v_pytest.log 2>&1; tail -15 gpurun_out/r02v_pytest.log | cut -c1-300
(time timeout 600 python bench.py) > gpurun_out/r02v_bench_default.json 2> gpurun_out/r02v_bench_default.err; tail -3 gpurun_out/r02v_bench_default.err
python - <<'PY'
import json
for l in open("gpurun_out/r02v_bench_default.json"):
    l = l.strip()
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], d["ms_steps"], "e2e", d["e2e"]["value"], "parity", d.get("parity_checked", {}).get("mismatches"), "cpu", d.get("cpu_baseline", {}).get("value"))
        for k, v in d.get("other_workloads", {}).items():
            print(" ", k, v.get("value"), v.get("skipped"), v.get("error"), (v.get("parity") or {}).get("mismatches"), (v.get("cpu_baseline") or {}).get("value"))
PY
