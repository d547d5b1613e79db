# round 2, call o: full GPU suite with --fades, then the round's ncu evidence
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/r02o_pytest.log 2>&1; tail -8 gpurun_out/r02o_pytest.log
X265CU_STAGE_MEMCPY=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02o_launches.csv python bench.py --no-cpu-baseline --no-others --no-e2e --steps 1 --warmup 0 --frames 120 > gpurun_out/r02o_ncu_launches.log 2>&1
tail -2 gpurun_out/r02o_launches.csv | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 6 -c 1 -o gpurun_out/prof_search_r02o python bench.py --no-cpu-baseline --no-others --no-e2e --steps 1 --warmup 0 --frames 200 > gpurun_out/r02o_ncu_search.log 2>&1
ls -la gpurun_out/*r02o*
