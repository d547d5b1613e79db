for i in 1 2 3 4; do
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/rep_$i.json 2> gpurun_out/rep_$i.err
done
