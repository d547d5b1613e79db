# what the driver runs at round end, in one go: GPU tests, smoke, the default bench line and the reference arm
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_final.log 2>&1; tail -4 gpurun_out/pytest_final.log
(time python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/smoke_final.log 2>&1; tail -3 gpurun_out/smoke_final.log
(time python bench.py) > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -4 gpurun_out/bench_final.err
(time python bench.py --impl reference) > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; tail -4 gpurun_out/bench_final_ref.err
