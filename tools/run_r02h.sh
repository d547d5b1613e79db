# round 2, call h: weightp decoupling + 4-lane search kernel.  Parity first (old kernel on the new pipeline, then everything
# on the new kernel), then A/B benches of the kernel / scheduling variants and one timeline.
mkdir -p gpurun_out
(X265CU_SEARCH_LANES=8 timeout 600 python -m pytest tests -m gpu -q -x -k "weights_assumed or fade or block_metrics or base8") > gpurun_out/r02h_pytest_lanes8.log 2>&1; tail -3 gpurun_out/r02h_pytest_lanes8.log
(time timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/r02h_pytest.log 2>&1; tail -15 gpurun_out/r02h_pytest.log
show() {
  python - "$1" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open("gpurun_out/r02h_%s.json" % n)); r = d["roofline"]
    print("%-14s value %.1f %s e2e %.1f us/job %.1f launches %d" % (n, d["value"], d["ms_steps"], d["e2e"]["value"], r["search_us_per_job"], r["search_launches_per_step"]))
    print("   kern", r["kernel_busy_ms_per_step"]); print("   host", r["host_ms_per_step"])
    if not d["e2e"].get("skipped"): print("   e2e ", d["e2e"].get("host_ms_last_step"))
except Exception as e:
    print(n, "failed", e)
PY
}
run() { # name, env..., -- bench args
  name=$1; shift
  env "$@" python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-others $EXTRA > gpurun_out/r02h_$name.json 2> gpurun_out/r02h_$name.err || tail -5 gpurun_out/r02h_$name.err
  show $name
}
EXTRA=""        run lanes8 X265CU_SEARCH_LANES=8
EXTRA=""        run lanes4 X265CU_SEARCH_LANES=4
EXTRA="--no-e2e" run lanes4_oneshot X265CU_SEARCH_ONESHOT=1
EXTRA="--no-e2e" run lanes8_oneshot X265CU_SEARCH_LANES=8 X265CU_SEARCH_ONESHOT=1
EXTRA="--no-e2e" run lanes4_mvploop X265CU_LIBDIR=$PWD/x265-amod_b200/lib_mvploop
EXTRA="--no-e2e" run lanes4_ctas24 X265CU_LIBDIR=$PWD/x265-amod_b200/lib_ctas24
EXTRA="--no-e2e" run lanes4_ctas16 X265CU_LIBDIR=$PWD/x265-amod_b200/lib_ctas16
EXTRA="--no-e2e --async-depth 32" run lanes4_a32 X265CU_SEARCH_LANES=4
EXTRA="--no-e2e --async-depth 64 --batch-min 16" run lanes4_a64_b16 X265CU_SEARCH_LANES=4
X265CU_TIMELINE=$PWD/gpurun_out/r02h_tl.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-others --no-e2e > gpurun_out/r02h_tl.json 2> gpurun_out/r02h_tl.err
python tools/timeline_summary.py gpurun_out/r02h_tl.csv 110 > gpurun_out/r02h_tl.txt 2>&1; cat gpurun_out/r02h_tl.txt
