# K4 code-size / occupancy variants against the default build: search time per job in a steady 200-frame run
mkdir -p gpurun_out
for v in "" sad3loop mvploop compact calls ctas24 ctas32; do
  if [ -z "$v" ]; then unset X265CU_LIBDIR; n=base; else export X265CU_LIBDIR=$PWD/x265-amod_b200/lib_$v; n=$v; fi
  python bench.py --steps 2 --warmup 1 --frames 200 --no-e2e --no-cpu-baseline > gpurun_out/var_$n.json 2> gpurun_out/var_$n.err || tail -3 gpurun_out/var_$n.err
  python - <<PY
import json
d=json.load(open("gpurun_out/var_$n.json")); r=d["roofline"]
print("%-9s value %.1f  search us/job %.1f busy %.0f ms  kernels %s" % ("$n", d["value"], r["search_us_per_job"], r["search_busy_ms_per_step"], r["kernel_busy_ms_per_step"]))
PY
done
