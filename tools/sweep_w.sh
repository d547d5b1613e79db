for bud in 24 20 28; do
X265CU_SEARCH_BUDGET=$bud timeout 300 python bench.py --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/sb_$bud.json 2> gpurun_out/sb_$bud.err
done
