# round 2, call i: compact search loop (two evaluation sites), 4-lane grouped cost kernel, per-slot mirror ordering,
# two-batches-in-flight scheduling.  Parity, then A/B benches, a timeline and one ncu capture of the search kernel.
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/r02i_pytest.log 2>&1; tail -12 gpurun_out/r02i_pytest.log
(X265CU_SEARCH_LANES=4 X265CU_COST_LANES=8 timeout 600 python -m pytest tests -m gpu -q -x -k "base8 or fade or pool16 or block_metrics or mc_metrics or b8 or slices or static") > gpurun_out/r02i_pytest_lanes4.log 2>&1; tail -4 gpurun_out/r02i_pytest_lanes4.log
show() {
  python - "$1" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open("gpurun_out/r02i_%s.json" % n)); r = d["roofline"]
    print("%-14s value %.1f %s e2e %.1f us/job %.1f launches %d" % (n, d["value"], d["ms_steps"], d["e2e"]["value"], r["search_us_per_job"], r["search_launches_per_step"]))
    print("   kern", r["kernel_busy_ms_per_step"]); print("   host", r["host_ms_per_step"])
    if not d["e2e"].get("skipped"): print("   e2e ", d["e2e"].get("host_ms_last_step"))
except Exception as e:
    print(n, "failed", e)
PY
}
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-others $EXTRA > gpurun_out/r02i_$name.json 2> gpurun_out/r02i_$name.err || tail -5 gpurun_out/r02i_$name.err
  show $name
}
EXTRA=""         run default X265CU_COST_LANES=4
EXTRA="--no-e2e" run cost8 X265CU_COST_LANES=8
EXTRA="--no-e2e" run inline X265CU_LIBDIR=$PWD/x265-amod_b200/lib_inline
EXTRA="--no-e2e" run compact4 X265CU_SEARCH_LANES=4
EXTRA="--no-e2e" run compact8_c24 X265CU_LIBDIR=$PWD/x265-amod_b200/lib_c24
EXTRA="--no-e2e" run compact4_c24 X265CU_SEARCH_LANES=4 X265CU_LIBDIR=$PWD/x265-amod_b200/lib_c4_24
EXTRA="--no-e2e" run oneshot X265CU_SEARCH_ONESHOT=1
EXTRA="--no-e2e --async-depth 96" run a96 X265CU_COST_LANES=4
X265CU_TIMELINE=$PWD/gpurun_out/r02i_tl.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-others --no-e2e > gpurun_out/r02i_tl.json 2> gpurun_out/r02i_tl.err
python tools/timeline_summary.py gpurun_out/r02i_tl.csv 110 > gpurun_out/r02i_tl.txt 2>&1; cat gpurun_out/r02i_tl.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 6 -c 1 -o gpurun_out/prof_search_r02i python bench.py --no-cpu-baseline --no-others --no-e2e --steps 1 --warmup 0 --frames 200 > gpurun_out/ncu_search_i.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cost_group_kernel4 -s 6 -c 1 -o gpurun_out/prof_cost4_r02i python bench.py --no-cpu-baseline --no-others --no-e2e --steps 1 --warmup 0 --frames 200 > gpurun_out/ncu_cost_i.log 2>&1
ls -la gpurun_out/*r02i*.ncu-rep
