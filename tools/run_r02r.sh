mkdir -p gpurun_out
(time timeout 500 python -m pytest tests -m gpu -q -k "mono400 or hist_scenecut or temporal_layers_2") > gpurun_out/r02r_pytest.log 2>&1; tail -25 gpurun_out/r02r_pytest.log | cut -c1-250
