# round 2, call l: job arrays / recalc results staged by a kernel instead of the copy engines; recycled timing events
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/r02l_pytest.log 2>&1; tail -6 gpurun_out/r02l_pytest.log
X265CU_HOST_TIMING=1 python bench.py --steps 4 --warmup 1 --no-cpu-baseline --no-others > gpurun_out/r02l_default.json 2> gpurun_out/r02l_default.err || tail -5 gpurun_out/r02l_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02l_default.json")); r = d["roofline"]
print("value %.1f %s e2e %.1f %s us/job %.1f" % (d["value"], d["ms_steps"], d["e2e"]["value"], d["e2e"].get("ms_steps"), r["search_us_per_job"]))
print("   kern", r["kernel_busy_ms_per_step"]); print("   host", r["host_ms_per_step"]); print("   e2e ", d["e2e"].get("host_ms_last_step"))
PY
grep "host timing" gpurun_out/r02l_default.err | tail -2
X265CU_TIMELINE=$PWD/gpurun_out/r02l_tl_e2e.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-others > gpurun_out/r02l_tl_e2e.json 2> gpurun_out/r02l_tl_e2e.err
python tools/timeline_summary.py gpurun_out/r02l_tl_e2e.csv 110 > gpurun_out/r02l_tl_e2e.txt 2>&1; cat gpurun_out/r02l_tl_e2e.txt
