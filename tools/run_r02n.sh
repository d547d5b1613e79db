# round 2, call n: the driver's own commands (default bench, reference arm) + the round's ncu evidence
mkdir -p gpurun_out
(time python bench.py) > gpurun_out/r02n_bench_default.json 2> gpurun_out/r02n_bench_default.err; tail -3 gpurun_out/r02n_bench_default.err
(time python bench.py --impl reference) > gpurun_out/r02n_bench_reference.json 2> gpurun_out/r02n_bench_reference.err; tail -3 gpurun_out/r02n_bench_reference.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02n_bench_default.json").read().strip().splitlines()[-1])
print("value", d["value"], d["ms_steps"], "e2e", d["e2e"]["value"], "parity", d.get("parity_checked", {}).get("mismatches"), "cpu", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("c_primitives_value"))
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "search_us_per_job", "search_jobs_per_step", "traffic")})
for k, v in d.get("other_workloads", {}).items():
    print(" ", k, json.dumps(v)[:400])
r = json.loads(open("gpurun_out/r02n_bench_reference.json").read().strip().splitlines()[-1])
print("reference", r.get("value"), r.get("cpu_baseline"))
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02n_launches.csv python bench.py --no-cpu-baseline --no-others --no-e2e --steps 1 --warmup 0 --frames 120 > gpurun_out/r02n_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 5 -c 1 -o gpurun_out/prof_search_r02n python bench.py --no-cpu-baseline --no-others --no-e2e --steps 1 --warmup 0 --frames 200 > gpurun_out/r02n_ncu_search.log 2>&1
ls -la gpurun_out/*r02n*
