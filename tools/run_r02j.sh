# round 2, call j: inline search + predictor loop as default, 4-lane grouped cost, mirrors through one scatter kernel.
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/r02j_pytest.log 2>&1; tail -6 gpurun_out/r02j_pytest.log
(X265CU_LIBDIR=$PWD/x265-amod_b200/lib_redux timeout 600 python -m pytest tests -m gpu -q -x -k "base8 or fade8 or block_metrics or mc_metrics or static") > gpurun_out/r02j_pytest_redux.log 2>&1; tail -3 gpurun_out/r02j_pytest_redux.log
show() {
  python - "$1" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open("gpurun_out/r02j_%s.json" % n)); r = d["roofline"]
    print("%-14s value %.1f %s e2e %.1f %s us/job %.1f launches %d" % (n, d["value"], d["ms_steps"], d["e2e"]["value"], d["e2e"].get("ms_steps"), r["search_us_per_job"], r["search_launches_per_step"]))
    print("   kern", r["kernel_busy_ms_per_step"]); print("   host", r["host_ms_per_step"])
    if not d["e2e"].get("skipped"): print("   e2e ", d["e2e"].get("host_ms_last_step"))
except Exception as e:
    print(n, "failed", e)
PY
}
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-others $EXTRA > gpurun_out/r02j_$name.json 2> gpurun_out/r02j_$name.err || tail -5 gpurun_out/r02j_$name.err
  show $name
}
EXTRA=""         run default X265CU_HOST_TIMING=1
grep "host timing" gpurun_out/r02j_default.err | tail -4
EXTRA=""         run memcpy_mirror X265CU_MIRROR_MEMCPY=1
EXTRA="--no-e2e" run nomvploop X265CU_LIBDIR=$PWD/x265-amod_b200/lib_nomvploop
EXTRA="--no-e2e" run redux X265CU_LIBDIR=$PWD/x265-amod_b200/lib_redux
EXTRA="--no-e2e" run oneshot X265CU_SEARCH_ONESHOT=1
EXTRA="--no-e2e" run green0 X265CU_GREEN=0
X265CU_TIMELINE=$PWD/gpurun_out/r02j_tl_e2e.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-others > gpurun_out/r02j_tl_e2e.json 2> gpurun_out/r02j_tl_e2e.err
python tools/timeline_summary.py gpurun_out/r02j_tl_e2e.csv 110 > gpurun_out/r02j_tl_e2e.txt 2>&1; cat gpurun_out/r02j_tl_e2e.txt
