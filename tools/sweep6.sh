timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r.log 2>&1; tail -3 gpurun_out/pytest_r.log
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/s6_calls_$i.json 2> gpurun_out/s6_calls_$i.err
done
cp tools/alt/libx265cu_inl.bin x265-amod_b200/lib/libx265cu.so
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/s6_inl_$i.json 2> gpurun_out/s6_inl_$i.err
done
