# round 2, call af: ncu capture of aq_edge_kernel (aq-mode 4) at 1080p
mkdir -p gpurun_out
X265CU_GREEN=0 X265CU_STAGE_MEMCPY=1 timeout 150 ncu --set full --clock-control none --import-source on -k regex:aq_edge_kernel -s 3 -c 1 -o gpurun_out/prof_aq_edge_r02af python -m pytest tests -m gpu -q -k "aq4_tl3" > gpurun_out/r02af_ncu.log 2>&1
tail -3 gpurun_out/r02af_ncu.log | cut -c1-200
ls -la gpurun_out/*r02af*
