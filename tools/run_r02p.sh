mkdir -p gpurun_out
CUDA_MODULE_LOADING=EAGER X265CU_STAGE_MEMCPY=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02p_launches_eager.csv python bench.py --no-cpu-baseline --no-others --no-e2e --steps 1 --warmup 0 --frames 120 > gpurun_out/r02p_a.log 2>&1
wc -l gpurun_out/r02p_launches_eager.csv; grep "^==ERROR" gpurun_out/r02p_launches_eager.csv | head -3
X265CU_GREEN=0 X265CU_STAGE_MEMCPY=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02p_launches_nogreen.csv python bench.py --no-cpu-baseline --no-others --no-e2e --steps 1 --warmup 0 --frames 120 > gpurun_out/r02p_b.log 2>&1
wc -l gpurun_out/r02p_launches_nogreen.csv; grep "^==ERROR" gpurun_out/r02p_launches_nogreen.csv | head -3
X265CU_GREEN=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02p_launches_nogreen_stage.csv python bench.py --no-cpu-baseline --no-others --no-e2e --steps 1 --warmup 0 --frames 120 > gpurun_out/r02p_c.log 2>&1
wc -l gpurun_out/r02p_launches_nogreen_stage.csv; grep "^==ERROR" gpurun_out/r02p_launches_nogreen_stage.csv | head -3
