# round 2, call k: e2e after the mirror scratch fix
mkdir -p gpurun_out
show() {
  python - "$1" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open("gpurun_out/r02k_%s.json" % n)); r = d["roofline"]
    print("%-14s value %.1f %s e2e %.1f %s us/job %.1f" % (n, d["value"], d["ms_steps"], d["e2e"]["value"], d["e2e"].get("ms_steps"), r["search_us_per_job"]))
    print("   host", r["host_ms_per_step"])
    if not d["e2e"].get("skipped"): print("   e2e ", d["e2e"].get("host_ms_last_step"))
except Exception as e:
    print(n, "failed", e)
PY
}
run() { name=$1; shift
  env "$@" python bench.py --steps 4 --warmup 1 --no-cpu-baseline --no-others $EXTRA > gpurun_out/r02k_$name.json 2> gpurun_out/r02k_$name.err || tail -5 gpurun_out/r02k_$name.err
  show $name
}
EXTRA="" run default X265CU_HOST_TIMING=1
grep "host timing" gpurun_out/r02k_default.err | tail -3
EXTRA="" run memcpy_mirror X265CU_MIRROR_MEMCPY=1
EXTRA="--async-depth 96" run a96 X265CU_HOST_TIMING=0
X265CU_TIMELINE=$PWD/gpurun_out/r02k_tl_e2e.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-others > gpurun_out/r02k_tl_e2e.json 2> gpurun_out/r02k_tl_e2e.err
python tools/timeline_summary.py gpurun_out/r02k_tl_e2e.csv 110 > gpurun_out/r02k_tl_e2e.txt 2>&1; cat gpurun_out/r02k_tl_e2e.txt
