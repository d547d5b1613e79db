mkdir -p gpurun_out
(time FUZZ_ENGINE=cuda timeout 36 python tools/fuzz_host_vs_reference.py seeds 150,254,443,476,723,1219,20,35,3,16,9003,9024,31,7,12,44,58,63,71,88,90,101,117,123,140) > gpurun_out/r02ah_fuzz_gpu.log 2>&1; grep -v "^\[\|^ \[" gpurun_out/r02ah_fuzz_gpu.log | cut -c1-300 | tail -12
