for cfg in "1 16 0" "1 16 34" "1 24 0" "1 8 0" "1 0 0" "1 16 12"; do set -- $cfg
X265CU_SEARCH_WORKERS=$3 timeout 300 python bench.py --no-cpu-baseline --steps 2 --warmup 1 --speculate $1 --async-depth $2 --pending-max 14 > gpurun_out/s3_$1_$2_$3.json 2> gpurun_out/s3_$1_$2_$3.err
done
