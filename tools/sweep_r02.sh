# scheduling sweep on the GPU box: async depth / batch size / workers per job against the K4 time per job
mkdir -p gpurun_out
for cfg in "32 0" "64 0" "64 16" "96 0" "96 24" "128 0"; do
  set -- $cfg
  python bench.py --steps 3 --warmup 2 --no-others --no-cpu-baseline --async-depth $1 --batch-min $2 > gpurun_out/sweep_$1_$2.json 2> gpurun_out/sweep_$1_$2.err || tail -3 gpurun_out/sweep_$1_$2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/sweep_$1_$2.json"))
r=d["roofline"]
print("async $1 batchMin $2: value %.1f e2e %.1f  search us/job %.1f busy %.0f ms launches %d jobs %d cost_jobs %d kernels %s" % (d["value"], d["e2e"]["value"], r["search_us_per_job"], r["search_busy_ms_per_step"], r["search_launches_per_step"], r["search_jobs_per_step"], r["cost_jobs_per_step"], r["kernel_busy_ms_per_step"]))
PY
done
