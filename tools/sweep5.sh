timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_p.log 2>&1; tail -3 gpurun_out/pytest_p.log
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --pending-max 16 > gpurun_out/s5_r64_$i.json 2> gpurun_out/s5_r64_$i.err
done
cp tools/alt/libx265cu_r72.bin x265-amod_b200/lib/libx265cu.so
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --pending-max 16 > gpurun_out/s5_r72_$i.json 2> gpurun_out/s5_r72_$i.err
done
