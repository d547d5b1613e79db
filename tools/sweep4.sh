timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_o.log 2>&1; tail -3 gpurun_out/pytest_o.log
for cfg in "1 16 0 8" "1 24 0 8" "1 32 0 12" "1 16 0 16" "1 16 24 8"; do set -- $cfg
X265CU_SEARCH_WORKERS=$3 timeout 300 python bench.py --no-cpu-baseline --steps 2 --warmup 1 --speculate $1 --async-depth $2 --pending-max $4 > gpurun_out/s4_$1_$2_$3_$4.json 2> gpurun_out/s4_$1_$2_$3_$4.err
done
