set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "golden or modes" > gpurun_out/pytest_k.log 2>&1; tail -3 gpurun_out/pytest_k.log
for w in 0 34 12 9; do for pm in 6 14; do
X265CU_SEARCH_WORKERS=$w timeout 300 python bench.py --no-cpu-baseline --steps 2 --warmup 1 --pending-max $pm > gpurun_out/sw_${w}_${pm}.json 2> gpurun_out/sw_${w}_${pm}.err
done; done
X265CU_SEARCH_WORKERS=0 timeout 300 python bench.py --no-cpu-baseline --steps 2 --warmup 1 --async-depth 0 > gpurun_out/sw_0_a0.json 2>&1
X265CU_SEARCH_WORKERS=0 timeout 300 python bench.py --no-cpu-baseline --steps 2 --warmup 1 --async-depth 32 --pending-max 14 > gpurun_out/sw_0_a32.json 2>&1
X265CU_SEARCH_WORKERS=0 timeout 300 python bench.py --no-cpu-baseline --steps 2 --warmup 1 --speculate 1 --async-depth 0 > gpurun_out/sw_0_s1.json 2>&1
