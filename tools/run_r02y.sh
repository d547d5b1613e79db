# round 2, call x: ncu capture of the --hme search kernel (one level-0 and one level-1 launch of a steady batch)
mkdir -p gpurun_out
X265CU_GREEN=0 X265CU_STAGE_MEMCPY=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:search_hme_kernel -s 0 -c 10 -o gpurun_out/prof_search_hme_r02y python bench.py --workload 1080p-8bit-hme --no-cpu-baseline --no-others --no-e2e --steps 1 --warmup 0 --frames 80 > gpurun_out/r02y_ncu_hme.log 2>&1
tail -3 gpurun_out/r02y_ncu_hme.log | cut -c1-300
ls -la gpurun_out/*r02y*
